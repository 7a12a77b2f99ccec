// ms_cyclic.cuh -- K2: flooding min-sum decoder for cyclic parity-check matrices, sm_100a.
//
// Replaces  min_sum__<Iterations,U,R,Q>(H, y, hor, vert)   reference codes/soft_decision.h:161-202
// (vertical__ :125-140, horizontal__ :101-122, column_sum :86-98, syndrome :79-84) for H built by
// cyclic::H<T>() (codes/cyclic.h:346-359): row r has its ones at columns r + tap[j].
//
// Mapping (one warp owns FPW frames; nothing but the final decisions/counters leaves the SM):
//   * lane <-> parity-check ROW.  The row's W edge messages r[j] live in REGISTERS for the whole
//     decode.  The shape of H (ms_shape.h: n, rows, tap offsets, rows per lane, frames per warp) is
//     a template parameter: every loop is fully unrolled and every shared-memory access is
//     [per-lane base register + immediate], so an edge costs no address arithmetic.
//   * per frame only y[n] and the column sums S[n] live in shared memory (2n floats).
//   * VN+CN pass: q_j = (S[c_j] - r_j) + y[c_j] exactly as soft_decision.h:135-136; min1/min2 and
//     the sign parity are reduced IN the lane (a row is private to a lane: no shuffles), then
//     r_j = +-fn_h(min over the others) (:106-118).
//   * column sums in the reference's order (rows ascending, :86-98): all lanes step through the
//     taps from the largest to the smallest and add r_j into S[row + tap_j]; within one step the
//     32 columns are distinct (conflict free), and a column receives its rows in ascending order
//     because row = column - tap.  One __syncwarp per step keeps that order.
//   * hard decision + stop test: lane <-> column, __ballot_sync gives the decided word as bit
//     masks, every row-lane popcounts its row mask against it (integer overlap mod 256 for the
//     reference's rule, parity for GF(2)).
//   * frames finish after different iteration counts: every frame group of a warp carries its own
//     iteration counter and pulls its next frame from a global queue head (atomicAdd) when done,
//     so no warp idles while others still iterate (persistent warps, dynamic schedule; results do
//     not depend on the schedule because the noise is keyed by the frame index).
//
// Float semantics: additions/multiplications are the explicit _rn intrinsics
// so nothing is contracted into FMAs; the OMS offset is applied in double; signum(0) = 0 is
// honoured because a zero among "the others" forces min = 0 and fn_h(0) = 0.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "channel.cuh"
#include "ms_params.h"
#include "ms_shape.h"

#ifndef CCGPU_MS_YREG_BUDGET
#define CCGPU_MS_YREG_BUDGET 28  /* messages + y values a lane keeps in registers (one-row-per-lane shapes), see YN; \
   measured on BCH(63,36), W = 18: y of 0 / 6 / 10 / 18 taps -> 2.25e8 / 2.28e8 / 2.30e8 / 2.26e8 frames/s */
#endif
#ifndef CCGPU_MS_YREG_RPL2
#define CCGPU_MS_YREG_RPL2 4  /* y values per row kept in registers by the two-rows-per-lane shapes: BCH(127,64) NMS 5 dB 0 -> 4.278e7 frames/s, 4 -> 4.324e7, 6 -> 4.260e7 */
#endif
#ifndef CCGPU_MS_CAP_W
#define CCGPU_MS_CAP_W 30  /* message registers per lane up to which the 64-register cap is applied */
#endif
#ifndef CCGPU_MS_MID_MINBLK
#define CCGPU_MS_MID_MINBLK 5  /* resident CTAs per SM asked for when a lane keeps 31..64 messages (102 registers): \
   measured +2 % BCH(127,64), +8 % (127,106), +24 % (127,99), +10 % (127,113) over the unconstrained allocation; 6 spills */
#endif
#ifndef CCGPU_MS_MINBLK
#define CCGPU_MS_MINBLK 8  /* resident CTAs per SM the small shapes are compiled for (64 registers) */
#endif
#ifndef CCGPU_MS_SMEM_COUNTERS
#define CCGPU_MS_SMEM_COUNTERS 1  /* the six per-warp statistics live in shared memory, not in registers */
#endif

#ifndef CCGPU_MS_PACK2
#define CCGPU_MS_PACK2 1  /* VN adds as packed add.rn.f32x2 (FADD2): two edges per issue slot, same IEEE results */
#endif
#ifndef CCGPU_MS_CN_PAIR
#define CCGPU_MS_CN_PAIR 1
#endif
#ifndef CCGPU_MS_UNIFORM
#define CCGPU_MS_UNIFORM 1  /* see UNI in the kernel: BCH(63,36) NMS 4 dB 2.303e8 -> 2.343e8 frames/s, 6 dB fused +7 %, (63,45) +3.6 % */
#endif
#ifndef CCGPU_MS_FNSCALE
#define CCGPU_MS_FNSCALE 1
#endif

#ifndef CCGPU_MS_VOLATILE_COLSUM
#define CCGPU_MS_VOLATILE_COLSUM 1  /* see VOLCS in the kernel */
#endif

namespace ccgpu {

constexpr unsigned kFull = 0xffffffffu;

// packed single-precision add/sub of sm_100 (add.rn.f32x2 -> FADD2): each half is the same IEEE
// round-to-nearest operation as __fadd_rn / __fsub_rn, so the results stay bit-identical
__device__ __forceinline__ void sub2_rn(float a0, float a1, float b0, float b1, float &d0, float &d1) {
  unsigned long long A, B, D;
  asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b0), "f"(b1));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(D) : "l"(A), "l"(B));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(D));
}
__device__ __forceinline__ void add2_rn(float a0, float a1, float b0, float b1, float &d0, float &d1) {
  unsigned long long A, B, D;
  asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(D) : "l"(A), "l"(B));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(D));
}

__device__ __forceinline__ float xor_sign(float v, unsigned signbits) {
  return __uint_as_float(__float_as_uint(v) ^ (signbits & 0x80000000u));
}
// value barrier: stops the compiler from re-associating an XOR that was folded once per row back
// into the per-edge path
__device__ __forceinline__ float opaque(float v) {
  asm volatile("" : "+f"(v));
  return v;
}

// fn_h of the variant applied to a non-negative minimum (soft_decision.h:204-213, :245-251)
__device__ __forceinline__ float cn_magnitude(const MsParams &p, float m) {
  if (p.variant == V_NMS || p.variant == V_NMS2D) return __fmul_rn(p.alpha_f, m);
  if (p.variant == V_OMS) {
    const double d = static_cast<double>(m) - p.beta_d;
    return static_cast<float>(d > 0.0 ? d : 0.0);
  }
  return m;
}
// the offset rule in double (:245-251), kept out of line so that the other variants do not issue its
// predicated-off FP64 sequence twice per row and iteration
static __device__ __noinline__ float2 cn_offset_pair(double beta, float m1, float m2) {
  const double d1 = static_cast<double>(m1) - beta, d2 = static_cast<double>(m2) - beta;
  return make_float2(static_cast<float>(d1 > 0.0 ? d1 : 0.0), static_cast<float>(d2 > 0.0 ? d2 : 0.0));
}
// fn_h applied to min1 and min2 of a row at once
__device__ __forceinline__ float2 cn_magnitude_pair(const MsParams &p, float m1, float m2) {
  if (p.variant == V_NMS || p.variant == V_NMS2D) return make_float2(__fmul_rn(p.alpha_f, m1), __fmul_rn(p.alpha_f, m2));
  if (p.variant == V_OMS) return cn_offset_pair(p.beta_d, m1, m2);
  return make_float2(m1, m2);
}

// guided self-scheduling: how many frame indices to take from the queue now -- work_batch while plenty are left,
// fewer towards the end so that no warp sits on a long private tail.  The queue head is not read (a load of the
// line every warp hammers with atomics costs a second round trip): it is estimated from the end of the warp's own
// previous batch plus what all warps took meanwhile (about warps x that batch; 4 x warps = 2^work_shift).
__device__ __forceinline__ unsigned guided_batch(const MsParams &p, long long own_next, unsigned last_batch) {
  const long long rem = static_cast<long long>(p.frames) - own_next - (static_cast<long long>(last_batch) << (p.work_shift - 2));
  const long long g = rem > 0 ? (rem >> p.work_shift) : 0;
  return g < 1 ? 1u : (g < static_cast<long long>(p.work_batch) ? static_cast<unsigned>(g) : p.work_batch);
}

__device__ __forceinline__ unsigned long long binom(unsigned n, unsigned r) {
  if (r > n) return 0ull;
  unsigned long long v = 1ull;
  for (unsigned i = 1; i <= r; ++i) v = v * (n - r + i) / i;
  return v;
}

// compact output layout (ccgpu_decode_llr_packed): 32 decided bits of a frame starting at bit `start` of the warp's
// concatenated decision words bw[0 .. NP) (start = colbase + 32 w for word w of the group's frame), `nbits` of them valid
template <int NP> __device__ __forceinline__ unsigned extract_word(const unsigned (&bw)[NP], int start, int nbits) {
  const int idx = start >> 5;
  unsigned lo = 0u, hi = 0u;
#pragma unroll
  for (int ps = 0; ps < NP; ++ps) {
    lo = (ps == idx) ? bw[ps] : lo;
    hi = (ps == idx + 1) ? bw[ps] : hi;
  }
  const unsigned v = __funnelshift_r(lo, hi, start & 31);
  return nbits >= 32 ? v : (v & ((1u << nbits) - 1u));
}

// resident CTAs per SM the register allocation aims at: 8 (64 registers) while the row's messages fit
template <class S, int VNQ> constexpr int ms_min_blocks() {
  constexpr int VN = VNQ >= VN_QUICK ? VNQ - VN_QUICK : VNQ;
  // measured: BCH(63,57), 32 messages per lane, 4.58e8 capped (spills) vs 4.82e8 free; self-correcting BCH(63,36),
  // 2 x 18 values per lane, 1.71e8 capped vs 1.42e8 free
  return ((VN == VN_SC || VN == VN_SPA) ? S::RPL * S::W <= 18 : S::RPL * S::W <= CCGPU_MS_CAP_W) ? CCGPU_MS_MINBLK
         : (VN == VN_PLAIN || VN == VN_2D) && S::RPL * S::W <= 64                                   ? CCGPU_MS_MID_MINBLK
                                                                                                     : 1;
}

template <class S, int VNQ>
__global__ void __launch_bounds__(kMsThreads, ms_min_blocks<S, VNQ>()) ms_cyclic_kernel(const __grid_constant__ MsParams p) {
  constexpr bool QUICK = VNQ >= VN_QUICK;                   // see quick_ok below
  constexpr int NBLK = (S::N + 3) >> 2;                     // Philox blocks (four channel values each) per frame
  constexpr int VN = QUICK ? VNQ - VN_QUICK : VNQ;
  constexpr int N = S::N, W = S::W, RPL = S::RPL, NP = S::NP, FPW = S::FPW;
  constexpr bool WRAP = S::WRAP, SC = VN == VN_SC, SPA = VN == VN_SPA;
  // one frame per warp: the schedule state (active, need_init, the iteration counter, the stop decision) is identical in
  // every lane by construction, so "does any lane ..." needs no vote
  constexpr bool UNI = CCGPU_MS_UNIFORM && FPW == 1;
  // fn_h as one multiply per minimum (see the check-node pass): measured BCH(63,36) NMS +1.0 %, MS +1.9 %, OMS +1.8 %, but
  // -0.4 .. -1.1 % on BCH(127,64) / (63,45) / (31,16): only the small one-frame-per-warp shapes use it
  constexpr bool FNS = CCGPU_MS_FNSCALE && FPW == 1 && RPL * W <= 18;
  // y of a row's edges is loop invariant: the first YN of them stay in registers (one shared-memory load less per
  // edge and iteration) as far as the 64-register budget of 8 resident CTAs per SM allows
  constexpr int YCAP = CCGPU_MS_YREG_BUDGET - W;
  constexpr int YN = (!SC && !SPA && !WRAP && RPL == 1 && YCAP > 0) ? ((YCAP < W ? YCAP : W) & ~1)
                     : (!SC && !SPA && !WRAP && RPL == 2 && FPW == 1) ? (CCGPU_MS_YREG_RPL2 & ~1) : 0;
  constexpr bool YREG = YN > 0;
  // ordered column sums: with one row per lane the read-modify-write chain goes through VOLATILE accesses, which
  // ptxas keeps in program order, instead of one __syncwarp per tap (ptxas proves the warp converged and turns
  // each of those into a NOP issue slot; measured +5 % on BCH(63,36)).  Two or more rows per lane keep the
  // __syncwarp form: volatile would serialise the rows of one step as well (measured -26 % on BCH(127,64)).
  constexpr bool VOLCS = CCGPU_MS_VOLATILE_COLSUM && RPL == 1;
  constexpr int ITEMS = FPW * N;            // columns handled by this warp, <= 32 * NP
  constexpr int SOFF = 32 * NP;             // S lives SOFF floats after y
  using T = typename S::taps;
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const int warp_in_cta = threadIdx.x >> 5;
  const int k = S::K > 0 ? S::K : p.k;      // rows (run time only for the redundant shapes)
  float *const ybuf = smem + warp_in_cta * (2 * SOFF);
  float *const sbuf = ybuf + SOFF;

  // ---------------- row-lane mapping
  int grp = 0;
  int row[RPL];
  bool rvalid[RPL];
  if (RPL == 1) {
    grp = (FPW > 1) ? lane / k : 0;
    row[0] = lane - grp * k;
    rvalid[0] = (FPW > 1) ? grp < FPW : lane < k;
    if (!rvalid[0]) grp = 0;  // idle lanes shadow group 0's control flow, touch nothing
  } else {
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      row[i] = lane + 32 * i;
      rvalid[i] = row[i] < k;
    }
  }
#pragma unroll
  for (int i = 0; i < RPL; ++i)
    if (!rvalid[i]) row[i] = 0;  // idle lanes compute on row 0's addresses and store nothing
  const int colbase = grp * N;
  const int lead_lane = grp * k;  // lane that speaks for the group (0 when FPW == 1)
  const bool is_lead = (lane == lead_lane) && rvalid[0];
  const unsigned gmask = (FPW > 1) ? (((1u << k) - 1u) << lead_lane) : kFull;
  // per-lane base pointers: edge (row, tap) is yrow[tap] / yrow[SOFF + tap]
  float *yrow[RPL];
#pragma unroll
  for (int i = 0; i < RPL; ++i) yrow[i] = ybuf + colbase + row[i];

  // ---------------- column-lane mapping: item c = lane + 32*pass  ->  (frame group, column)
  int cgrp_lead[NP];   // lead lane of the group that owns item c
  int ccol[NP];        // column inside the frame
  bool cvalid[NP];
  unsigned cmask[NP];  // bits of the concatenated decision word that belong to MY group
#pragma unroll
  for (int ps = 0; ps < NP; ++ps) {
    const int c = lane + 32 * ps;
    cvalid[ps] = c < ITEMS;
    const int f = (FPW > 1 && cvalid[ps]) ? c / N : 0;
    ccol[ps] = c - f * N;
    cgrp_lead[ps] = f * k;
    const int lo = colbase - 32 * ps, hi = colbase + N - 32 * ps;
    unsigned m = 0;
    if (hi > 0 && lo < 32) {
      const int a = lo < 0 ? 0 : lo, b = hi > 32 ? 32 : hi;
      m = (b - a >= 32) ? kFull : (((1u << (b - a)) - 1u) << a);
    }
    cmask[ps] = m;
  }

  // ---------------- row masks for the stop test (bit = colbase + column)
  unsigned rmask[RPL][NP];
#pragma unroll
  for (int i = 0; i < RPL; ++i) {
#pragma unroll
    for (int ps = 0; ps < NP; ++ps) rmask[i][ps] = 0;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      int c = row[i] + T::get(j);
      if (WRAP && c >= N) c -= N;
      c += colbase;
#pragma unroll
      for (int ps = 0; ps < NP; ++ps)
        if ((c >> 5) == ps && rvalid[i]) rmask[i][ps] |= 1u << (c & 31);
    }
  }

  // ---------------- per-group decode state (replicated in every lane of the group)
  const long long units = static_cast<long long>(gridDim.x) * (kMsThreads / 32) * FPW;
  long long my_frame = (static_cast<long long>(blockIdx.x) * (kMsThreads / 32) + warp_in_cta) * FPW + grp;
  bool active = my_frame < static_cast<long long>(p.frames);
  bool need_init = true;
  int it = 0;
  float r[RPL][W];
  float qold[(SC || SPA) ? RPL : 1][(SC || SPA) ? W : 1];  // previous q (SCMS) / prefix products (SPA)
  float yreg[RPL][YN > 0 ? YN : 1];
  // 32-bit per-warp counters (a warp sees far fewer than 2^32 / 50 frames per launch); widened at the end
#if CCGPU_MS_SMEM_COUNTERS
  // every thread owns one slot per statistic (only lead lanes ever write): six registers freed for the decoder
  __shared__ unsigned cnt_s[6][kMsThreads];
#pragma unroll
  for (int s = 0; s < 6; ++s) cnt_s[s][threadIdx.x] = 0u;
#else
  unsigned cnt_frames = 0, cnt_ferr = 0, cnt_berr = 0, cnt_iter = 0, cnt_fail = 0, cnt_und = 0;
#endif
  static_assert(FPW <= 8, "grant slots");
  __shared__ long long pool_next_s[kMsThreads / 32];  // frame indices already taken from the queue: next one ..
  __shared__ int pool_left_s[kMsThreads / 32];        // .. and how many are left
  __shared__ long long grant_s[kMsThreads / 32][FPW > 1 ? FPW : 1];
  __shared__ unsigned pool_batch_s[kMsThreads / 32];   // size of the batch taken last
  if (lane == 0) {
    pool_left_s[warp_in_cta] = 0;
    pool_batch_s[warp_in_cta] = 0;
    pool_next_s[warp_in_cta] = units;  // the queue starts behind the statically assigned first frames
  }
  __syncwarp();

  // dynamic schedule: every finishing group (`done`, warp-uniform when FPW == 1) gets the next frame index of the
  // warp's pool, which is refilled from the global queue head with ONE atomic per batch of up to work_batch frames
  // (one atomic per frame on one address capped the small codes); returns the index in every lane of the group
  auto take_frames = [&](bool done) -> long long {
    long long next = 0;
    if (FPW == 1) {
      if (done && lane == 0) {
        int left = pool_left_s[warp_in_cta];
        next = pool_next_s[warp_in_cta];
        if (left == 0) {
          left = static_cast<int>(guided_batch(p, next, pool_batch_s[warp_in_cta]));
          pool_batch_s[warp_in_cta] = static_cast<unsigned>(left);
          next = units + static_cast<long long>(atomicAdd(p.work, static_cast<unsigned long long>(left)));
        }
        pool_next_s[warp_in_cta] = next + 1;
        pool_left_s[warp_in_cta] = left - 1;
      }
    } else {
      const unsigned leadm = __ballot_sync(kFull, done && is_lead);
      if (lane == 0) {
        int left = pool_left_s[warp_in_cta], slot = 0;
        long long nx = pool_next_s[warp_in_cta];
        for (unsigned m = leadm; m; m &= m - 1u, ++slot) {
          if (left == 0) {
            left = static_cast<int>(guided_batch(p, nx, pool_batch_s[warp_in_cta]));
            pool_batch_s[warp_in_cta] = static_cast<unsigned>(left);
            nx = units + static_cast<long long>(atomicAdd(p.work, static_cast<unsigned long long>(left)));
          }
          grant_s[warp_in_cta][slot] = nx++;
          --left;
        }
        pool_next_s[warp_in_cta] = nx;
        pool_left_s[warp_in_cta] = left;
      }
      __syncwarp();
      if (done && is_lead) next = grant_s[warp_in_cta][__popc(leadm & ((1u << lane) - 1u))];
      __syncwarp();  // grant_s may be rewritten by the next call
    }
    return __shfl_sync(kFull, next, lead_lane);
  };

  // QUICK flavour: a frame whose channel values are all positive is decided by iteration 0 without executing it --
  // q = y > 0 on every edge, so every check-node message is >= 0, every total L_c = S_c + y_c > 0, the decided word is
  // all-zero and both stop rules hold (soft_decision.h:161-202 with r = 0, S = 0).  The outputs are the same (bits 0,
  // iteration index 0, no failure, one iteration counted); only the totals L would need the column sums, so the
  // shortcut is off when L is asked for.  At high Eb/N0 -- where a sweep spends most of its frames -- this is the
  // common case (59 % of the BCH(63,36) frames at 7 dB: 1.6e9 -> 2.0e9 frames/s).  It is a separate instantiation
  // because the extra branch costs the plain kernel 3 % at 4 dB (register allocation at the 64-register cap); the
  // host picks it when enough such frames are expected (api.cu launch_ms).
  const bool quick_ok = QUICK && p.L == nullptr && p.stop_rule != STOP_NONE && p.max_iter >= 1;
  // SCREENING (QUICK flavour, one frame per warp, Monte-Carlo points that want counters only): the warp draws frames in
  // passes of FP = 32 / NBLK -- every lane generates ONE Philox block, so all lanes work (16 of 32 did for n = 63) --
  // parks the values in a staging row and looks at the signs.  All-positive frames are counted and forgotten: no y / S
  // initialisation, no message reset, no trip through the iteration loop, no queue bookkeeping (about 230 warp
  // instructions per such frame before, about 60 now).  The frames that need the decoder are then taken from the staging
  // row one after the other.  Same counters as any other schedule: the noise is keyed by the frame index.
  constexpr bool SCREEN = QUICK && FPW == 1 && NBLK <= 32;
  constexpr int FP = SCREEN ? 32 / NBLK : 1;  // frames per screening pass
  const bool screen_on = SCREEN && quick_ok && p.src == SRC_PHILOX && p.counters != nullptr && p.bits == nullptr && p.iter == nullptr &&
                         p.failed == nullptr && p.packed == nullptr && p.status == nullptr;
  __shared__ __align__(16) float stage_s[SCREEN ? kMsThreads / 32 : 1][SCREEN ? 128 : 4];  // FP frames of 4 * NBLK values
  __shared__ long long stage_idx_s[SCREEN ? kMsThreads / 32 : 1][FP];
  unsigned hardmask = 0;   // staged frames that still need the decoder
  bool first_pass = true;  // the first staged frame of a warp is its statically assigned one

#if CCGPU_MS_FNSCALE
  const bool is_oms = p.variant == V_OMS;
  const float fn_scale = (p.variant == V_NMS || p.variant == V_NMS2D) ? p.alpha_f : 1.0f;
#endif
  while (true) {
    if (UNI ? !active : (__ballot_sync(kFull, active) == 0u)) break;
    bool skip = false;

    // ============ (re)fill frame groups that finished
    const unsigned initm = UNI ? ((active && need_init) ? kFull : 0u) : __ballot_sync(kFull, active && need_init);
    if (SCREEN && screen_on) {
      if (active && need_init) {  // warp-uniform: one frame per warp
        __syncwarp();
        for (;;) {
          if (hardmask == 0u) {  // ---- a new screening pass
            if (lane == 0) {
              int left = pool_left_s[warp_in_cta];
              long long nx = pool_next_s[warp_in_cta];
#pragma unroll
              for (int f = 0; f < FP; ++f) {
                if (first_pass && f == 0) {
                  stage_idx_s[warp_in_cta][0] = my_frame;
                  continue;
                }
                if (left == 0) {
                  left = static_cast<int>(guided_batch(p, nx, pool_batch_s[warp_in_cta]));
                  pool_batch_s[warp_in_cta] = static_cast<unsigned>(left);
                  nx = units + static_cast<long long>(atomicAdd(p.work, static_cast<unsigned long long>(left)));
                }
                stage_idx_s[warp_in_cta][f] = nx++;
                --left;
              }
              pool_next_s[warp_in_cta] = nx;
              pool_left_s[warp_in_cta] = left;
            }
            first_pass = false;
            __syncwarp();
            const int f = lane / NBLK, blk = lane - f * NBLK;
            const long long fr = f < FP ? stage_idx_s[warp_in_cta][f] : static_cast<long long>(p.frames);
            bool nonpos = false;
            if (fr < static_cast<long long>(p.frames)) {
              const float4 v = awgn_block(p.keys, p.point, p.frame0 + static_cast<uint64_t>(fr), blk, p.sigma);
              const float4 sv = make_float4(v.x * p.llr_scale, v.y * p.llr_scale, v.z * p.llr_scale, v.w * p.llr_scale);
              nonpos = !(sv.x > 0.0f) || (4 * blk + 1 < N && !(sv.y > 0.0f)) || (4 * blk + 2 < N && !(sv.z > 0.0f)) ||
                       (4 * blk + 3 < N && !(sv.w > 0.0f));
              *reinterpret_cast<float4 *>(&stage_s[warp_in_cta][4 * lane]) = sv;
            }
            const unsigned npm = __ballot_sync(kFull, nonpos);
            unsigned nquick = 0, nvalid = 0;
#pragma unroll
            for (int g = 0; g < FP; ++g) {
              if (stage_idx_s[warp_in_cta][g] < static_cast<long long>(p.frames)) {
                ++nvalid;
                constexpr unsigned gm = NBLK >= 32 ? kFull : ((1u << NBLK) - 1u);
                if (npm & (gm << (g * NBLK))) hardmask |= 1u << g;
                else ++nquick;
              }
            }
            if (lane == 0 && nquick) {  // iteration 0 decides them: one iteration each, no errors
#if CCGPU_MS_SMEM_COUNTERS
              cnt_s[0][threadIdx.x] += nquick;
              cnt_s[3][threadIdx.x] += nquick;
#else
              cnt_frames += nquick;
              cnt_iter += nquick;
#endif
            }
            if (nvalid == 0u) {  // the queue is exhausted
              active = false;
              break;
            }
            __syncwarp();
            if (hardmask == 0u) continue;
          }
          // ---- the next staged frame that needs the decoder
          const int f = __ffs(hardmask) - 1;
          hardmask &= hardmask - 1u;
          my_frame = stage_idx_s[warp_in_cta][f];
#pragma unroll
          for (int ps = 0; ps < NP; ++ps) {
            const int c = lane + 32 * ps;
            if (c < N) {
              ybuf[c] = stage_s[warp_in_cta][f * 4 * NBLK + c];
              sbuf[c] = 0.0f;
            }
          }
#pragma unroll
          for (int i = 0; i < RPL; ++i)
#pragma unroll
            for (int j = 0; j < W; ++j) {
              r[i][j] = 0.0f;
              if (SC || SPA) qold[i][j] = 0.0f;
            }
          it = 0;
          need_init = false;
          break;
        }
        __syncwarp();
        if (YREG && active) {
#pragma unroll
          for (int i = 0; i < RPL; ++i)
#pragma unroll
            for (int j = 0; j < YN; ++j) yreg[i][j] = yrow[i][T::get(j)];
        }
      }
      if (!active) continue;  // exhausted: the loop head ends the warp
    } else if (initm) {
      __syncwarp();  // the finished frame's y / S were read by other lanes (outputs): order those reads first
      if (p.src == SRC_HBM) {
#pragma unroll
        for (int ps = 0; ps < NP; ++ps) {
          const long long fr = __shfl_sync(kFull, my_frame, cgrp_lead[ps]);
          if (cvalid[ps] && ((initm >> cgrp_lead[ps]) & 1u)) {
            ybuf[lane + 32 * ps] = __ldg(p.y + fr * N + ccol[ps]);
            sbuf[lane + 32 * ps] = 0.0f;
          }
        }
      } else if (p.src == SRC_PHILOX) {
#pragma unroll
        for (int b0 = 0; b0 < FPW * NBLK; b0 += 32) {
          const int b = b0 + lane;
          const bool bv = b < FPW * NBLK;
          const int f = (FPW > 1 && bv) ? b / NBLK : 0;
          const int blk = b - f * NBLK;
          const int src_lane = f * k;
          const long long fr = __shfl_sync(kFull, my_frame, src_lane);
          if (bv && ((initm >> src_lane) & 1u)) {
            const float4 v = awgn_block(p.keys, p.point, p.frame0 + static_cast<uint64_t>(fr), blk, p.sigma);
            const int c0 = f * N + 4 * blk;
            const float vv[4] = { v.x * p.llr_scale, v.y * p.llr_scale, v.z * p.llr_scale, v.w * p.llr_scale };
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (4 * blk + e < N) {
                ybuf[c0 + e] = vv[e];
                sbuf[c0 + e] = 0.0f;
              }
          }
        }
      } else {  // SRC_BITFLIP: x = -2*bit + 1 for the pattern of lexicographic rank frame0 + fr
        if (is_lead && active && need_init) {
          unsigned long long rank = p.frame0 + static_cast<unsigned long long>(my_frame);
          unsigned ones = p.flip_weight;
          for (int c = 0; c < N; ++c) {
            const unsigned long long zero_first = binom(N - c - 1, ones);
            float v = 1.0f;
            if (rank >= zero_first && ones > 0) {
              rank -= zero_first;
              --ones;
              v = -1.0f;
            }
            ybuf[colbase + c] = v;
            sbuf[colbase + c] = 0.0f;
          }
        }
      }
      if (active && need_init) {
#pragma unroll
        for (int i = 0; i < RPL; ++i)
#pragma unroll
          for (int j = 0; j < W; ++j) {
            r[i][j] = 0.0f;
            if (SC || SPA) qold[i][j] = 0.0f;
          }
        it = 0;
        need_init = false;
      }
      __syncwarp();
      if (QUICK && quick_ok) {
        unsigned nonpos = 0;
#pragma unroll
        for (int ps = 0; ps < NP; ++ps) {
          const bool np = cvalid[ps] && !(ybuf[lane + 32 * ps] > 0.0f);
          nonpos |= __ballot_sync(kFull, np) & cmask[ps];
        }
        const bool quick = active && it == 0 && ((initm >> lead_lane) & 1u) && nonpos == 0u;  // my group: fresh, all y > 0
        skip = __all_sync(kFull, !active || quick);  // every frame of this warp: otherwise the body runs for all of them
      }
      if (YREG && !(QUICK && skip)) {
#pragma unroll
        for (int i = 0; i < RPL; ++i)
#pragma unroll
          for (int j = 0; j < YN; ++j) yreg[i][j] = yrow[i][T::get(j)];
      }
    }

    unsigned bw[NP];
    bool stop;
    if (QUICK && skip) {
#pragma unroll
      for (int ps = 0; ps < NP; ++ps) bw[ps] = 0u;
      stop = true;
    } else {
    // ============ VN + CN  (vertical__ / horizontal__)
    if (SPA) {
      // extension (not in the reference): sum-product / tanh rule.  r_j = 2 atanh( prod_{i != j} tanh(q_i / 2) ),
      // the exclusive product as prefix * suffix, clamped like oracle/ms_oracle.c so that atanh stays finite
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        float prod = 1.0f;
#pragma unroll
        for (int j = 0; j < W; ++j) {
          int off = T::get(j);
          if (WRAP && row[i] + off >= N) off -= N;
          const float q = __fadd_rn(__fsub_rn(yrow[i][SOFF + off], r[i][j]), yrow[i][off]);
          const float th = tanhf(0.5f * q);
          qold[i][j] = prod;  // product of the taps before j
          prod *= th;
          r[i][j] = th;
        }
        float suffix = 1.0f;  // product of the taps after j
#pragma unroll
        for (int j = W - 1; j >= 0; --j) {
          const float th = r[i][j];
          float pr = qold[i][j] * suffix;
          pr = fminf(fmaxf(pr, -0.99999994f), 0.99999994f);
          r[i][j] = 2.0f * atanhf(pr);
          suffix *= th;
        }
      }
    } else {
    float f1s[RPL], f2s[RPL], m1v[RPL];
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      float m1 = FLT_MAX, m2 = FLT_MAX;
      unsigned par = 0;
      float qv[W];
      // ---- q_j = (S[c_j] - r_j) + y[c_j]  (:135-136), two edges per packed instruction
      constexpr int WP = CCGPU_MS_PACK2 ? (W & ~1) : 0;
#pragma unroll
      for (int j = 0; j < WP; j += 2) {
        int off0 = T::get(j), off1 = T::get(j + 1);
        if (WRAP && row[i] + off0 >= N) off0 -= N;
        if (WRAP && row[i] + off1 >= N) off1 -= N;
        float e0, e1;
        sub2_rn(yrow[i][SOFF + off0], yrow[i][SOFF + off1], r[i][j], r[i][j + 1], e0, e1);
        if (VN == VN_2D) {  // normalised_vertical (:215-218); scalar on purpose: ptxas contracts a packed
          e0 = __fmul_rn(p.beta_f, e0);  // mul.rn.f32x2 + add.rn.f32x2 pair into FFMA2 (one rounding), which
          e1 = __fmul_rn(p.beta_f, e1);  // the reference does not do
        }
        add2_rn(e0, e1, j < YN ? yreg[i][j < YN ? j : 0] : yrow[i][off0], j + 1 < YN ? yreg[i][j + 1 < YN ? j + 1 : 0] : yrow[i][off1],
                qv[j], qv[j + 1]);
      }
#pragma unroll
      for (int j = WP; j < W; ++j) {
        int off = T::get(j);
        if (WRAP && row[i] + off >= N) off -= N;
        float e = __fsub_rn(yrow[i][SOFF + off], r[i][j]);
        if (VN == VN_2D) e = __fmul_rn(p.beta_f, e);
        qv[j] = __fadd_rn(e, j < YN ? yreg[i][j < YN ? j : 0] : yrow[i][off]);
      }
#pragma unroll
      for (int j = 0; j < W; ++j) {
        float q = qv[j];
        if (SC) {
          const float qo = qold[i][j];
          if (p.variant == V_SCMS1) {                          // :261-267
            const bool keep = (qo == 0.0f) || ((qo > 0.0f) == (q > 0.0f) && (qo < 0.0f) == (q < 0.0f));
            q = keep ? q : 0.0f;
          } else {                                             // SCMS2 :275-281
            q = (__fmul_rn(q, qo) > 0.0f) ? q : __fmul_rn(0.5f, __fadd_rn(q, qo));
          }
          qold[i][j] = q;
        } else {
          r[i][j] = q;  // r_j is dead once q_j exists; reuse its register
        }
        const float a = fabsf(q);
        m2 = fminf(m2, fmaxf(m1, a));
        m1 = fminf(m1, a);
        par ^= __float_as_uint(q);
      }
      m1v[i] = m1;
#if CCGPU_MS_FNSCALE
      // (compiled in for every shape, used where measured faster: FNS)
      // fn_h as one multiply per minimum: m1 / m2 are never NaN (fminf keeps the other operand), so x 1.0f is exact for
      // the unnormalised variants; the offset rule (double intermediate) is a rare warp-uniform call
      float2 g;
      if (FNS) {
        g = make_float2(__fmul_rn(fn_scale, m1), __fmul_rn(fn_scale, m2));
        if (is_oms) g = cn_offset_pair(p.beta_d, m1, m2);
      } else {
        g = cn_magnitude_pair(p, m1, m2);
      }
#elif CCGPU_MS_CN_PAIR
      const float2 g = cn_magnitude_pair(p, m1, m2);
#else
      const float2 g = make_float2(cn_magnitude(p, m1), cn_magnitude(p, m2));
#endif
      f1s[i] = opaque(xor_sign(g.x, par));  // fold the row's sign parity in once
      f2s[i] = opaque(xor_sign(g.y, par));
    }
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const float q = SC ? qold[i][j] : r[i][j];
        // min over the others is min2 exactly for the edge(s) attaining min1 (ties: min2 == min1)
        const float f = (fabsf(q) == m1v[i]) ? f2s[i] : f1s[i];
        r[i][j] = xor_sign(f, __float_as_uint(q));  // sign = prod of the other signs (:114,:118)
      }
    }
    }
    __syncwarp();

    // ============ column sums, rows ascending (column_sum :86-98)
#pragma unroll
    for (int ps = 0; ps < NP; ++ps)
      if (cvalid[ps]) sbuf[lane + 32 * ps] = 0.0f;
    __syncwarp();
#pragma unroll
    for (int pass = 0; pass < (WRAP ? 2 : 1); ++pass) {
#pragma unroll
      for (int j = W - 1; j >= 0; --j) {
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
          int off = T::get(j);
          bool wrapped = false;
          if (WRAP && row[i] + off >= N) {
            off -= N;
            wrapped = true;
          }
          if (rvalid[i] && wrapped == (pass == 1)) {
            if (VOLCS) {
              volatile float *sp = yrow[i] + SOFF + off;
              *sp = __fadd_rn(*sp, r[i][j]);
            } else {
              yrow[i][SOFF + off] = __fadd_rn(yrow[i][SOFF + off], r[i][j]);
            }
          }
        }
        if (!VOLCS) __syncwarp();  // without it ptxas lets the next step's loads overtake this step's stores (measured)
      }
    }
    if (VOLCS) __syncwarp();

    // ============ totals, hard decision (:178-183), stop test (:79-84)
#pragma unroll
    for (int ps = 0; ps < NP; ++ps) {
      bool neg = false;
      if (cvalid[ps]) neg = __fadd_rn(sbuf[lane + 32 * ps], ybuf[lane + 32 * ps]) < 0.0f;
      bw[ps] = __ballot_sync(kFull, neg);
    }
    if (p.stop_simple) {
      // reference rule on a matrix whose rows cover every column with weight < 256: every overlap is
      // zero exactly when the decided word is all-zero (host sets the flag, see api.cu fill_decoder)
      unsigned anyone = 0;
#pragma unroll
      for (int ps = 0; ps < NP; ++ps) anyone |= bw[ps] & cmask[ps];
      stop = anyone == 0u;
    } else {
      bool bad = false;
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        int ov = 0;
#pragma unroll
        for (int ps = 0; ps < NP; ++ps) ov += __popc(bw[ps] & rmask[i][ps]);
        if (rvalid[i]) {
          if (p.stop_rule == STOP_REF) bad |= (ov & 255) != 0;
          else if (p.stop_rule == STOP_GF2) bad |= (ov & 1) != 0;
          else bad = true;
        }
      }
      const unsigned badm = __ballot_sync(kFull, bad);
      stop = (badm & gmask) == 0u;
    }
    }  // !(QUICK && skip)
    const bool last = it + 1 >= p.max_iter;
    const bool fin = active && (stop || last);
    const unsigned finm = UNI ? (fin ? kFull : 0u) : __ballot_sync(kFull, fin);
    if (finm) {
      // ---- decided word / totals of the finishing groups
      if (p.bits != nullptr || p.L != nullptr) {
#pragma unroll
        for (int ps = 0; ps < NP; ++ps) {
          const long long fr = __shfl_sync(kFull, my_frame, cgrp_lead[ps]);
          if (cvalid[ps] && ((finm >> cgrp_lead[ps]) & 1u)) {
            if (p.bits) p.bits[fr * N + ccol[ps]] = static_cast<uint8_t>((bw[ps] >> lane) & 1u);
            if (p.L) p.L[fr * N + ccol[ps]] = __fadd_rn(sbuf[lane + 32 * ps], ybuf[lane + 32 * ps]);
          }
        }
      }
      long long next = 0;
      if (fin) {
        const bool failed = !stop && p.stop_rule != STOP_NONE;
        int nbits = 0;
#pragma unroll
        for (int ps = 0; ps < NP; ++ps) nbits += __popc(bw[ps] & cmask[ps]);
        if (is_lead) {
          if (p.iter) p.iter[my_frame] = static_cast<uint8_t>(failed ? p.max_iter : it);
          if (p.failed) p.failed[my_frame] = failed ? 1 : 0;
          if (p.packed) {
            constexpr int NPW = (N + 31) >> 5;
#pragma unroll
            for (int w = 0; w < NPW; ++w) p.packed[my_frame * NPW + w] = extract_word<NP>(bw, colbase + 32 * w, N - 32 * w);
          }
          if (p.status) p.status[my_frame] = static_cast<uint8_t>(failed ? 255 : it);
#if CCGPU_MS_SMEM_COUNTERS
          cnt_s[0][threadIdx.x] += 1u;
          cnt_s[3][threadIdx.x] += static_cast<unsigned>(it + 1);
          if (failed || nbits != 0) {  // the error statistics: rare at the Eb/N0 where most frames are simulated
            cnt_s[1][threadIdx.x] += 1u;
            cnt_s[2][threadIdx.x] += static_cast<unsigned>(nbits);
            cnt_s[4][threadIdx.x] += failed ? 1u : 0u;
            cnt_s[5][threadIdx.x] += failed ? 0u : 1u;
          }
#else
          cnt_frames += 1;
          cnt_iter += static_cast<unsigned>(it + 1);
          cnt_fail += failed ? 1 : 0;
          cnt_berr += static_cast<unsigned>(nbits);
          cnt_ferr += (failed || nbits != 0) ? 1 : 0;
          cnt_und += (!failed && nbits != 0) ? 1 : 0;
#endif
        }
      }
      if (SCREEN && screen_on) {
        if (fin) need_init = true;  // the refill above finds the warp's next frame
      } else {
        next = take_frames(fin);
        if (fin) {
          my_frame = next;
          active = my_frame < static_cast<long long>(p.frames);
          need_init = true;
        }
      }
    }
    if (!fin) ++it;
  }

  // ---------------- counters: warp reduce, one atomic per slot per warp
  if (p.counters != nullptr) {
#if CCGPU_MS_SMEM_COUNTERS
    unsigned long long v[6];
#pragma unroll
    for (int s = 0; s < 6; ++s) v[s] = cnt_s[s][threadIdx.x];
#else
    unsigned long long v[6] = { cnt_frames, cnt_ferr, cnt_berr, cnt_iter, cnt_fail, cnt_und };  // widen
#endif
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      unsigned long long x = v[s];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
      if (lane == 0 && x) atomicAdd(p.counters + s, x);
    }
  }
}

}  // namespace ccgpu
