// ms_cyclic_q.cuh -- K2q: FIXED-POINT flooding min-sum for cyclic parity-check matrices, sm_100a.
//
// Extension (north_star "fixed-point min-sum"; no reference implementation): min_sum__ of the reference's
// codes/soft_decision.h:161-202 (vertical__ :125-140, horizontal__ :101-122, column_sum :86-98, syndrome :79-84)
// with Q = R = integer, as restated in oracle/ms_oracle.c (oracle_min_sum_fixed) and specified in include/ccgpu.h
// (CCGPU_MS_Q / NMS_Q / OMS_Q).  Results are bit-identical to that restatement.
//
// Mapping: as in ms_cyclic.cuh (lane <-> parity-check row, messages in registers, frame data in shared memory,
// persistent warps with a dynamic frame queue), with two differences that integer arithmetic allows:
//   * TWO FRAMES PER LANE.  Every 32-bit register / shared-memory word carries the same quantity of two
//     independent frames ("slots") in its 16-bit halves, and every instruction of the iteration works on both:
//     one issue slot, one shared-memory wavefront per TWO edge updates.  The integers travel as fp16x2 (HADD2 /
//     HMNMX2 / HFMA2): every value of the decoder is an integer of magnitude <= 2048, which fp16 represents and
//     adds EXACTLY, and the fp16 pipe gives |x| and -x as free operand modifiers (the int16x2 DPX forms need three
//     instructions for an absolute value; measured pipe rates in tools/ubench/pipes.cu).  The host rejects
//     parameter sets whose column sums could leave that range (api.cu check_params_q).
//     A slot that finishes (stop test or iteration limit) is refilled on its own; the other slot keeps iterating.
//   * integer addition is associative: the channel value is folded into the column accumulator
//     (S'_c = y_c + sum_rows r, so q = S' - r and L = S': one load and one add per edge less than the float
//     kernel), and the order of the column sum does not matter (only its read-modify-write hazard is kept in
//     program order, as in ms_cyclic.cuh).
//   * the "min over the others" select is arithmetic: t = (|q| > min1) as 1.0 / 0.0 per half (one HSET2.BF) is 0 exactly
//     on the edge(s) that attain min1, r = f2 + (f1 - f2) * t (one HFMA2, exact on integers): per-half predicates do
//     not exist, and a mask + two LOP3 would all sit on the alu pipe, the busiest one.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "channel.cuh"
#include "ms_cyclic.cuh"
#include "ms_params.h"
#include "ms_shape.h"

namespace ccgpu {

// ---- packed fp16x2 helpers on raw 32-bit words (both halves always carry small integers)
__device__ __forceinline__ unsigned h2_sub(unsigned a, unsigned b) {
  unsigned d;
  asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ unsigned h2_add(unsigned a, unsigned b) {
  unsigned d;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ unsigned h2_abs(unsigned a) {
  unsigned d;
  asm("abs.f16x2 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ unsigned h2_min(unsigned a, unsigned b) {
  unsigned d;
  asm("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ unsigned h2_max(unsigned a, unsigned b) {
  unsigned d;
  asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ unsigned h2_fma(unsigned a, unsigned b, unsigned c) {
  unsigned d;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ unsigned h2_neg(unsigned a) {
  unsigned d;
  asm("neg.f16x2 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
// per half: |a| > b ? 1.0 : 0.0  (HSET2.BF.GT with the |a| operand modifier)
__device__ __forceinline__ unsigned h2_gt_abs(unsigned a, unsigned b) {
  unsigned d;
  asm("{.reg .b32 t; abs.f16x2 t, %1; set.gt.f16x2.f16x2 %0, t, %2;}" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ unsigned short int_to_h(int v) { return __half_as_ushort(__int2half_rn(v)); }
__device__ __forceinline__ unsigned h2_splat(int v) { return 0x10001u * int_to_h(v); }

// quantiser of the fixed-point decoder: clamp(rint(y * scale), +-ymax), NaN -> 0 (oracle_quantise in ms_oracle.c)
__device__ __forceinline__ unsigned short quantise_h(float y, float scale, int ymax) {
  const float t = __fmul_rn(y, scale);
  int v = __float2int_rn(t);  // NaN -> 0, +-inf / large -> INT_MAX / INT_MIN, ties to even
  v = max(-ymax, min(ymax, v));
  return int_to_h(v);  // 0 -> +0
}

#ifndef CCGPU_Q_HSET
#define CCGPU_Q_HSET 1  /* argmin select through HSET2.BF (1) or min(|q| - min1, 1) (0) */
#endif
#ifndef CCGPU_Q_QUICK
#define CCGPU_Q_QUICK 1  /* compile the all-positive shortcut into the refill loop */
#endif
#ifndef CCGPU_Q_MINBLK
#define CCGPU_Q_MINBLK 8  /* resident CTAs per SM the small shapes (<= 22 packed messages per lane) are compiled for */
#endif
template <class S> constexpr int ms_q_min_blocks() {
  // packed messages per lane -> resident CTAs per SM the register allocation aims at (64 / 102 / 168 registers)
  return S::RPL * S::W <= 22 ? CCGPU_Q_MINBLK : S::RPL * S::W <= 40 ? 5 : S::RPL * S::W <= 72 ? 3 : 1;
}

template <class S>
__global__ void __launch_bounds__(kMsThreads, ms_q_min_blocks<S>()) ms_cyclic_q_kernel(const __grid_constant__ MsParams p) {
  constexpr int N = S::N, W = S::W, RPL = S::RPL, NP = S::NP, FPW = S::FPW;
  constexpr bool WRAP = S::WRAP;
  constexpr bool VOLCS = CCGPU_MS_VOLATILE_COLSUM && RPL == 1;  // see ms_cyclic.cuh
  constexpr int ITEMS = FPW * N;
  constexpr int SOFF = 32 * NP;
  constexpr unsigned SIGN2 = 0x80008000u;
  using T = typename S::taps;
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const int warp_in_cta = threadIdx.x >> 5;
  const int k = S::K > 0 ? S::K : p.k;
  unsigned *const ybuf = reinterpret_cast<unsigned *>(smem) + warp_in_cta * (2 * SOFF);  // y of both slots, packed
  unsigned *const sbuf = ybuf + SOFF;                                                   // S' = y + sum r

  // ---------------- row-lane mapping (ms_cyclic.cuh)
  int grp = 0;
  int row[RPL];
  bool rvalid[RPL];
  if (RPL == 1) {
    grp = (FPW > 1) ? lane / k : 0;
    row[0] = lane - grp * k;
    rvalid[0] = (FPW > 1) ? grp < FPW : lane < k;
    if (!rvalid[0]) grp = 0;
  } else {
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      row[i] = lane + 32 * i;
      rvalid[i] = row[i] < k;
    }
  }
#pragma unroll
  for (int i = 0; i < RPL; ++i)
    if (!rvalid[i]) row[i] = 0;
  const int colbase = grp * N;
  const int lead_lane = grp * k;
  const bool is_lead = (lane == lead_lane) && rvalid[0];
  const unsigned gmask = (FPW > 1) ? (((1u << k) - 1u) << lead_lane) : kFull;
  unsigned *yrow[RPL];
#pragma unroll
  for (int i = 0; i < RPL; ++i) yrow[i] = ybuf + colbase + row[i];

  // ---------------- column-lane mapping
  int cgrp_lead[NP];
  int ccol[NP];
  bool cvalid[NP];
  unsigned cmask[NP];
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

  // ---------------- row masks for the general stop test
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

  // ---------------- per-(group, slot) decode state, replicated in every lane of the group
  const long long units = static_cast<long long>(gridDim.x) * (kMsThreads / 32) * FPW * 2;
  long long my_frame[2];
  bool active[2], need_init[2];
  int it[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    my_frame[h] = ((static_cast<long long>(blockIdx.x) * (kMsThreads / 32) + warp_in_cta) * FPW + grp) * 2 + h;
    active[h] = my_frame[h] < static_cast<long long>(p.frames);
    need_init[h] = true;
    it[h] = 0;
  }
  unsigned r[RPL][W];  // check-node messages of both slots (fp16x2)
#pragma unroll
  for (int i = 0; i < RPL; ++i)
#pragma unroll
    for (int j = 0; j < W; ++j) r[i][j] = 0u;
  __shared__ unsigned cnt_s[6][kMsThreads];
#pragma unroll
  for (int s = 0; s < 6; ++s) cnt_s[s][threadIdx.x] = 0u;
  constexpr int NBLK = (N + 3) >> 2;
  static_assert(FPW <= 8, "grant slots");
  __shared__ long long pool_next_s[kMsThreads / 32];
  __shared__ int pool_left_s[kMsThreads / 32];
  __shared__ long long grant_s[kMsThreads / 32][FPW > 1 ? FPW : 1];
  __shared__ unsigned pool_batch_s[kMsThreads / 32];
  if (lane == 0) {
    pool_left_s[warp_in_cta] = 0;
    pool_batch_s[warp_in_cta] = 0;
    pool_next_s[warp_in_cta] = units;
  }
  // unused lanes / slots compute on whatever is in shared memory: make it finite once
#pragma unroll
  for (int ps = 0; ps < NP; ++ps) {
    ybuf[lane + 32 * ps] = 0u;
    sbuf[lane + 32 * ps] = 0u;
  }
  __syncwarp();

  // frame queue: identical to ms_cyclic.cuh (one atomic per batch of indices, guided batch size)
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
      __syncwarp();
    }
    return __shfl_sync(kFull, next, lead_lane);
  };

  // constants of the check-node function, splat over both slots:
  //   g = max(rne(A m / 1024) - B, 0) with (A, B) = (1024, 0) MS_Q, (A, 0) NMS_Q, (1024, B) OMS_Q.
  // fma(m, A/1024, 1024) rounds the exact sum ONCE to the fp16 grid, whose spacing is 1 in [1024, 2048): that is
  // rne(A m / 1024) + 1024 (m <= q_msg_max <= 1023, A <= 1024; A/1024 is a fp16 number)
  const unsigned kMmax = p.q_h2_mmax, kAlpha = p.q_h2_alpha, k1024 = p.q_h2_1024, k1024B = p.q_h2_1024b;  // fp16x2, host-made
  unsigned short *const ybuf16 = reinterpret_cast<unsigned short *>(ybuf);
  unsigned short *const sbuf16 = reinterpret_cast<unsigned short *>(sbuf);
  // the all-positive shortcut needs no totals (L is only known after the column sums) and a stop rule
  const bool quick_ok = CCGPU_Q_QUICK && p.quick_hint != 0 && p.L == nullptr && p.stop_rule != STOP_NONE && p.max_iter >= 1;

  while (true) {
    if (__ballot_sync(kFull, active[0] || active[1]) == 0u) break;

    // ============ (re)fill the slots that finished.  Both slots of a lane are served by ONE pass (for BCH(63,36) the 2 x 16
    // Philox blocks of two frames keep all 32 lanes busy).  A fresh frame whose quantised channel values are all
    // positive is decided by iteration 0: q = y > 0 on every edge, so every check-node message is >= 0, every total
    // L = y + sum r > 0, the decided word is all-zero and both stop rules hold (soft_decision.h:161-202 with r = 0).
    // When every active slot of the warp holds such a frame the iteration is skipped -- same outputs (bits 0,
    // iteration index 0, no failure, one iteration counted).  At high Eb/N0, where a sweep spends most of its frames,
    // that is the common case (a retirement loop inside the refill measured 15 % slower at 4 dB: code growth).
    bool skip = false;
    while (true) {  // (one pass; `break` leaves it)
      bool need[2];
      unsigned initm[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        need[h] = active[h] && need_init[h];
        initm[h] = __ballot_sync(kFull, need[h]);
      }
      if ((initm[0] | initm[1]) == 0u) break;
      __syncwarp();  // the finished frame's totals were read by other lanes (outputs)
      if (p.src == SRC_HBM) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (initm[h] == 0u) continue;
#pragma unroll
          for (int ps = 0; ps < NP; ++ps) {
            const long long fr = __shfl_sync(kFull, my_frame[h], cgrp_lead[ps]);
            if (cvalid[ps] && ((initm[h] >> cgrp_lead[ps]) & 1u)) {
              const unsigned short v = quantise_h(__ldg(p.y + fr * N + ccol[ps]), p.q_scale, p.q_ymax);
              ybuf16[2 * (lane + 32 * ps) + h] = v;
              sbuf16[2 * (lane + 32 * ps) + h] = v;
            }
          }
        }
      } else if (p.src == SRC_PHILOX) {
        constexpr int PER_SLOT = FPW * NBLK;  // Philox blocks of one slot of this warp
#pragma unroll
        for (int b0 = 0; b0 < 2 * PER_SLOT; b0 += 32) {
          const int b = b0 + lane;
          const bool bv = b < 2 * PER_SLOT;
          const int h = (bv && b >= PER_SLOT) ? 1 : 0;
          const int bb = b - h * PER_SLOT;
          const int f = (FPW > 1 && bv) ? bb / NBLK : 0;
          const int blk = bb - f * NBLK;
          const int src_lane = f * k;
          const long long fr0 = __shfl_sync(kFull, my_frame[0], src_lane), fr1 = __shfl_sync(kFull, my_frame[1], src_lane);
          if (bv && (((h ? initm[1] : initm[0]) >> src_lane) & 1u)) {
            const float4 v = awgn_block(p.keys, p.point, p.frame0 + static_cast<uint64_t>(h ? fr1 : fr0), blk, p.sigma);
            const int c0 = f * N + 4 * blk;
            const float vv[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (4 * blk + e < N) {
                const unsigned short q = quantise_h(vv[e], p.q_scale, p.q_ymax);
                ybuf16[2 * (c0 + e) + h] = q;
                sbuf16[2 * (c0 + e) + h] = q;
              }
          }
        }
      } else {  // SRC_BITFLIP
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (is_lead && need[h]) {
            unsigned long long rank = p.frame0 + static_cast<unsigned long long>(my_frame[h]);
            unsigned ones = p.flip_weight;
            const unsigned short plus = quantise_h(1.0f, p.q_scale, p.q_ymax), minus = quantise_h(-1.0f, p.q_scale, p.q_ymax);
            for (int c = 0; c < N; ++c) {
              const unsigned long long zero_first = binom(N - c - 1, ones);
              unsigned short v = plus;
              if (rank >= zero_first && ones > 0) {
                rank -= zero_first;
                --ones;
                v = minus;
              }
              ybuf16[2 * (colbase + c) + h] = v;
              sbuf16[2 * (colbase + c) + h] = v;
            }
          }
        }
      }
      {
        const unsigned keep = (need[0] ? 0u : 0x0000ffffu) | (need[1] ? 0u : 0xffff0000u);  // zero the fresh slots' messages (+0)
#pragma unroll
        for (int i = 0; i < RPL; ++i)
#pragma unroll
          for (int j = 0; j < W; ++j) r[i][j] &= keep;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (need[h]) {
          it[h] = 0;
          need_init[h] = false;
        }
      __syncwarp();
      // ---- all-positive fresh frames: if every active slot of this warp holds one, the iteration is not executed
      if (quick_ok) {
        bool mine_quick = true;  // my group's active slots are fresh and all-positive
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          unsigned nonpos = 0;
          if (initm[h] != 0u) {
#pragma unroll
            for (int ps = 0; ps < NP; ++ps) {
              const short v = static_cast<short>(ybuf16[2 * (lane + 32 * ps) + h]);  // fp16 pattern: <= 0 as an integer iff the value is <= 0
              nonpos |= __ballot_sync(kFull, cvalid[ps] && v <= 0) & cmask[ps];
            }
          }
          if (active[h] && !(need[h] && nonpos == 0u)) mine_quick = false;
        }
        skip = __all_sync(kFull, mine_quick);
      }
      break;
    }
    unsigned bw[2][NP];
    bool fin[2], stopv[2];
    if (skip) {
#pragma unroll
      for (int ps = 0; ps < NP; ++ps) bw[0][ps] = bw[1][ps] = 0u;
      stopv[0] = stopv[1] = true;
    } else {
    // ============ VN + CN for both slots at once (vertical__ / horizontal__)
    unsigned f2s[RPL], ds[RPL], m1v[RPL];
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      unsigned m1 = 0x7bff7bffu, m2 = 0x7bff7bffu;  // largest finite fp16 = "numeric_limits::max()" (:107)
      unsigned par = 0;
#pragma unroll
      for (int j = 0; j < W; ++j) {
        int off = T::get(j);
        if (WRAP && row[i] + off >= N) off -= N;
        const unsigned q = h2_sub(yrow[i][SOFF + off], r[i][j]);  // q = (S - r) + y with y folded into S' (:135-136)
        r[i][j] = q;
        const unsigned a = h2_abs(q);
        m2 = h2_min(m2, h2_max(m1, a));
        m1 = h2_min(m1, a);
        par ^= q;
      }
      // r = sign * fn_h(min(min, q_msg_max)) (:118); min over the others = min2 on the edge(s) attaining min1
      const unsigned g1 = h2_max(h2_sub(h2_fma(h2_min(m1, kMmax), kAlpha, k1024), k1024B), 0u);
      const unsigned g2 = h2_max(h2_sub(h2_fma(h2_min(m2, kMmax), kAlpha, k1024), k1024B), 0u);
      const unsigned f1 = g1 ^ (par & SIGN2);  // fold the row's sign parity in once
      const unsigned f2 = g2 ^ (par & SIGN2);
      f2s[i] = f2;
      ds[i] = h2_sub(f1, f2);   // f1 - f2: r = f2 + (f1 - f2) * t
      m1v[i] = m1;
    }
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const unsigned q = r[i][j];
#if CCGPU_Q_HSET
        const unsigned t = h2_gt_abs(q, m1v[i]);  // 0.0 on the edge(s) attaining min1, else 1.0
#else
        const unsigned t = h2_min(h2_sub(h2_abs(q), m1v[i]), 0x3c003c00u);
#endif
        const unsigned f = h2_fma(ds[i], t, f2s[i]);
        r[i][j] = f ^ (q & SIGN2);  // times the sign of the edge's own q: product of the OTHER signs (:114,:118)
      }
    }
    __syncwarp();

    // ============ column accumulators S' = y + sum_rows r (column_sum :86-98; the order is irrelevant for integers)
#pragma unroll
    for (int ps = 0; ps < NP; ++ps)
      if (cvalid[ps]) sbuf[lane + 32 * ps] = ybuf[lane + 32 * ps];
    __syncwarp();
#pragma unroll
    for (int j = W - 1; j >= 0; --j) {
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        int off = T::get(j);
        if (WRAP && row[i] + off >= N) off -= N;
        if (rvalid[i]) {
          if (VOLCS) {
            volatile unsigned *sp = yrow[i] + SOFF + off;
            *sp = h2_add(*sp, r[i][j]);
          } else {
            yrow[i][SOFF + off] = h2_add(yrow[i][SOFF + off], r[i][j]);
          }
        }
      }
      if (!VOLCS) __syncwarp();
    }
    if (VOLCS) __syncwarp();

    // ============ totals L = S', hard decision (:178-183), stop test (:79-84), per slot
#pragma unroll
    for (int ps = 0; ps < NP; ++ps) {
      // L < 0 is the sign bit: a total is never -0 (y enters as +0, x - x = +0, +0 + -0 = +0 in round-to-nearest)
      const unsigned x = cvalid[ps] ? sbuf[lane + 32 * ps] : 0u;
      bw[0][ps] = __ballot_sync(kFull, (x & 0x8000u) != 0u);
      bw[1][ps] = __ballot_sync(kFull, (x & 0x80000000u) != 0u);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      bool stop;
      if (p.stop_simple) {
        unsigned anyone = 0;
#pragma unroll
        for (int ps = 0; ps < NP; ++ps) anyone |= bw[h][ps] & cmask[ps];
        stop = anyone == 0u;
      } else {
        bool bad = false;
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
          int ov = 0;
#pragma unroll
          for (int ps = 0; ps < NP; ++ps) ov += __popc(bw[h][ps] & rmask[i][ps]);
          if (rvalid[i]) {
            if (p.stop_rule == STOP_REF) bad |= (ov & 255) != 0;
            else if (p.stop_rule == STOP_GF2) bad |= (ov & 1) != 0;
            else bad = true;
          }
        }
        const unsigned badm = __ballot_sync(kFull, bad);
        stop = (badm & gmask) == 0u;
      }
      stopv[h] = stop;
    }
    }  // !skip
#pragma unroll
    for (int h = 0; h < 2; ++h) fin[h] = active[h] && (stopv[h] || it[h] + 1 >= p.max_iter);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const unsigned finm = __ballot_sync(kFull, fin[h]);
      if (finm) {
        if (p.bits != nullptr || p.L != nullptr) {
#pragma unroll
          for (int ps = 0; ps < NP; ++ps) {
            const long long fr = __shfl_sync(kFull, my_frame[h], cgrp_lead[ps]);
            if (cvalid[ps] && ((finm >> cgrp_lead[ps]) & 1u)) {
              if (p.bits) p.bits[fr * N + ccol[ps]] = static_cast<uint8_t>((bw[h][ps] >> lane) & 1u);
              if (p.L) p.L[fr * N + ccol[ps]] = __half2float(__ushort_as_half(sbuf16[2 * (lane + 32 * ps) + h]));
            }
          }
        }
        if (fin[h]) {
          const bool failed = !stopv[h] && p.stop_rule != STOP_NONE;
          int nbits = 0;
#pragma unroll
          for (int ps = 0; ps < NP; ++ps) nbits += __popc(bw[h][ps] & cmask[ps]);
          if (is_lead) {
            if (p.iter) p.iter[my_frame[h]] = static_cast<uint8_t>(failed ? p.max_iter : it[h]);
            if (p.failed) p.failed[my_frame[h]] = failed ? 1 : 0;
            if (p.packed) {
              constexpr int NPW = (N + 31) >> 5;
#pragma unroll
              for (int w = 0; w < NPW; ++w)
                p.packed[my_frame[h] * NPW + w] = extract_word<NP>(bw[h], colbase + 32 * w, N - 32 * w);
            }
            if (p.status) p.status[my_frame[h]] = static_cast<uint8_t>(failed ? 255 : it[h]);
            cnt_s[0][threadIdx.x] += 1u;
            cnt_s[3][threadIdx.x] += static_cast<unsigned>(it[h] + 1);
            if (failed || nbits != 0) {
              cnt_s[1][threadIdx.x] += 1u;
              cnt_s[2][threadIdx.x] += static_cast<unsigned>(nbits);
              cnt_s[4][threadIdx.x] += failed ? 1u : 0u;
              cnt_s[5][threadIdx.x] += failed ? 0u : 1u;
            }
          }
        }
        const long long next = take_frames(fin[h]);
        if (fin[h]) {
          my_frame[h] = next;
          active[h] = my_frame[h] < static_cast<long long>(p.frames);
          need_init[h] = true;
        }
      }
      if (!fin[h]) ++it[h];
    }
  }

  // ---------------- counters
  if (p.counters != nullptr) {
    unsigned long long v[6];
#pragma unroll
    for (int s = 0; s < 6; ++s) v[s] = cnt_s[s][threadIdx.x];
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
