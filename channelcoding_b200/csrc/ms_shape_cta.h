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
};

}  // namespace ccgpu
