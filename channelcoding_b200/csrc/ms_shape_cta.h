// ms_shape_cta.h -- compile-time description of a cyclic parity-check matrix for ms_cyclic_cta.cuh.
#pragma once
#include "ms_shape.h"

namespace ccgpu {

// N columns; K rows (0 = run time); RPL rows per thread; WPF warps per frame (= per CTA)
template <int N_, int K_, int RPL_, int WPF_, bool WRAP_, class TAPS> struct ShapeCta {
  static constexpr int N = N_, K = K_, RPL = RPL_, WPF = WPF_, W = TAPS::count;
  static constexpr int THREADS = 32 * WPF_;
  static constexpr int NPW = (N_ + 31) / 32;                       // 32-bit words of the decided word
  static constexpr int CPASS = (N_ + THREADS - 1) / THREADS;      // column passes
  static constexpr bool WRAP = WRAP_;
  using taps = TAPS;
  static_assert(K_ <= THREADS * RPL_, "rows must fit the threads");
  // gather form of the ordered column sum (ms_cyclic_cta.cuh): message (row, tap j) is parked at X[j][row]; a tap's
  // array is XS floats apart from the next one: THREADS rows and a 32-float guard band that is never written
  static constexpr int TMAX = TAPS::get(TAPS::count - 1);
  static constexpr int XS = THREADS + 32;
  static constexpr int XBYTES = (32 + W * XS) * 4;
  static constexpr bool GATHER = RPL_ == 1 && XBYTES <= 56 * 1024;
  // y / S entries: with wrap-around in gather form the first TMAX columns are mirrored behind column N - 1
  static constexpr int YW = (GATHER && WRAP_) ? ((N_ + TMAX + 31) / 32) * 32 : NPW * 32;
  static constexpr int DYN_SMEM = GATHER ? XBYTES : 0;
};

}  // namespace ccgpu
