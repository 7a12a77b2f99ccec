// ms_shape.h -- compile-time description of a cyclic parity-check matrix for ms_cyclic.cuh.
#pragma once

namespace ccgpu {

// column offsets of the ones of row 0, ascending; folded to immediates after unrolling
template <int... T> struct Taps {
  static constexpr int count = sizeof...(T);
  __host__ __device__ static constexpr int get(int j) {
    constexpr int a[] = { T... };
    return a[j];
  }
};

// N columns; K rows (0 = given at run time, redundant H); RPL rows per lane; FPW frames per warp;
// NP 32-column passes (FPW * N <= 32 * NP); WRAP: row r has ones at (r + tap) mod N
template <int N_, int K_, int RPL_, int FPW_, int NP_, bool WRAP_, class TAPS> struct Shape {
  static constexpr int N = N_, K = K_, RPL = RPL_, FPW = FPW_, NP = NP_, W = TAPS::count;
  static constexpr bool WRAP = WRAP_;
  using taps = TAPS;
  static_assert(FPW_ * N_ <= 32 * NP_, "columns of the warp's frames must fit the passes");
  static_assert(K_ == 0 || (K_ <= 32 * RPL_ && (RPL_ > 1 || FPW_ * K_ <= 32)), "rows must fit the lanes");
  static_assert(RPL_ == 1 || FPW_ == 1, "several frames per warp only with one row per lane");
};

}  // namespace ccgpu
