// ms_cyclic_cta.cuh -- K2c: the cyclic min-sum decoder of ms_cyclic.cuh for codes whose rows do not
// fit one warp with register-resident messages (BCH(255,131): 124 rows of weight 68).
//
// Same arithmetic, same order, same outputs as ms_cyclic_kernel (reference codes/soft_decision.h:161-202);
// the difference is the mapping: one CTA of WPF warps owns ONE frame, thread <-> parity-check row(s)
// (row = tid + THREADS * i), the W messages of a row stay in registers, y[n] / S[n] / the decided
// word (as 32-bit masks) live in shared memory.  CTAs are persistent and pull frames from the same
// global queue head.
//
// Ordered column sums (column_sum, soft_decision.h:86-98: a column adds its rows in ascending order, and float
// addition does not commute with that order), two forms:
//   * GATHER (ShapeCta::GATHER, one row per thread): every row thread parks its new message of tap j at X[j][row]
//     in shared memory (one store per edge, immediate offset).  After ONE block barrier the thread that owns
//     column c adds X[W-1][c - t_{W-1}], ..., X[0][c - t_0] in a register (then X[.][c + N - t_j] for the wrapped
//     edges of a redundant H): descending taps are ascending rows.  Rows that do not exist are never written and
//     read +0.0 from the start of the kernel, and S + (+0.0) == S bit for bit because a running sum that started
//     at +0.0 is never -0.0.  Whether a (column, tap) pair has a row at all is decided PER WARP AT COMPILE TIME
//     (the column loop is instantiated once per warp index: the 32 columns of a warp and the tap offset are
//     constants), so only pairs some lane needs are loaded -- 1.2 loads per edge for BCH(255,131) instead of the
//     2.06 of a dense sweep; the lanes of such a warp that have no row read a zero guard band (32 floats between the
//     taps' arrays).  All loads are independent; three block barriers per iteration.
//     With wrap-around, y and S carry a copy of their first TMAX entries behind entry N - 1, so that the row
//     phase addresses edge (row, tap) at [row + tap] without a wrap test.
//   * scatter (the shapes whose X does not fit): read-modify-write of S[row + tap_j], the steps separated by
//     __syncthreads (W per iteration; the kernel is then bound by the barrier latency).
//
// Monte-Carlo points at high Eb/N0 (SRC_PHILOX, counters only): frames are taken in GROUPS of WPF, every warp
// generates the channel values of its own frame of the group (all lanes busy: N / 128 Philox blocks per lane) and tests
// them; a frame whose values are all positive is decided by iteration 0 (quick_ok in ms_cyclic.cuh) and is only
// counted -- no shared memory, no block barrier.  The frames of the group that need the decoder are then decoded one
// after the other by the whole CTA from the warps' staging rows.  At 11 dB (98 % of the BCH(255,131) frames are
// all-positive) the frame-at-a-time form spent its time in seven block barriers per frame with half the threads idle
// during the noise generation; results are identical because the noise is keyed by the frame index.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "channel.cuh"
#include "ms_cyclic.cuh"
#include "ms_params.h"
#include "ms_shape.h"
#include "ms_shape_cta.h"

#ifndef CCGPU_MS_CTA_MINBLK
#define CCGPU_MS_CTA_MINBLK 4  /* one row per thread, plain / 2-D flavour: 4 CTAs per SM = 128 registers, no spills; \
   measured: 5 (96 registers, 108 bytes spilled) is 7 % slower, an unconstrained allocation (168 registers, 3 CTAs) 12 % slower */
#endif

#ifndef CCGPU_MS_CTA_SMALL_W
#define CCGPU_MS_CTA_SMALL_W 32  /* gather shapes with at most this many messages per thread ... */
#endif
#ifndef CCGPU_MS_CTA_SMALL_MINBLK
#define CCGPU_MS_CTA_SMALL_MINBLK 6  /* ... are compiled for this many CTAs per SM */
#endif

#ifndef CCGPU_MS_CTA_YN
#define CCGPU_MS_CTA_YN 16  /* y values of a row the small gather shapes keep in registers.  Measured on the 127-row \
   BCH(127,64), NMS 5 dB: 0 -> 2.037e7 frames/s, 8 -> 2.119e7, 16 -> 2.185e7 (79 registers, no spills), 22 -> 2.078e7 \
   (spills), 30 -> 1.953e7 */
#endif
#ifndef CCGPU_MS_CTA_BIG_YN
#define CCGPU_MS_CTA_BIG_YN 24  /* the same for the shapes with more than CCGPU_MS_CTA_SMALL_W messages per thread. \
   Measured on BCH(255,131), NMS 4 dB, four CTAs per SM (128 registers): 0 -> 3.704e6 frames/s, 8 -> 3.69e6, 16 -> 3.73e6, \
   24 -> 3.80e6 (32 bytes spilled), 32 -> 3.07e6 (104 bytes spilled); three CTAs per SM (168 registers): 32 -> 3.73e6, \
   48 -> 3.74e6, 68 -> 3.20e6 */
#endif

namespace ccgpu {

// the two-rows-per-thread (wrap-around) shapes and the self-correcting flavour need more than 128 registers
template <class S, int VN> constexpr int ms_cta_min_blocks() {
  if (S::RPL == 1 && VN != VN_SC && S::GATHER && S::W <= CCGPU_MS_CTA_SMALL_W) return CCGPU_MS_CTA_SMALL_MINBLK;
  return (S::RPL == 1 && VN != VN_SC) ? CCGPU_MS_CTA_MINBLK : 1;
}


// GATHER column phase of warp WARP (compile time): the warp's columns are c = 32 WARP + THREADS cp + lane.  A
// (column pass, tap) pair is loaded only if some lane of the warp has a row for it; lanes that have none read the
// guard band (zero).  Writes S[c] (and its mirror) and the decision masks.
template <class S, int WARP>
__device__ __forceinline__ void cta_gather_columns(int tid, int k, const float *xs, const float *ybuf, float *sbuf, unsigned *bword) {
  constexpr int N = S::N, W = S::W, THREADS = S::THREADS, CPASS = S::CPASS, XS = S::XS;
  constexpr bool WRAP = S::WRAP;
  constexpr int KMAX = S::K > 0 ? S::K : (THREADS < N ? THREADS : N);  // rows that can exist
  using T = typename S::taps;
  const int lane = tid & 31;
#pragma unroll
  for (int cp = 0; cp < CPASS; ++cp) {
    const int cmin = 32 * WARP + THREADS * cp, cmax = cmin + 31;  // constants after unrolling
    const int c = cmin + lane;
    const float *xc = xs + c;
    float acc = 0.0f;
#pragma unroll
    for (int j = W - 1; j >= 0; --j)  // rows c - t_j, ascending as j descends
      if (cmax - T::get(j) >= 0 && cmin - T::get(j) <= KMAX - 1) acc = __fadd_rn(acc, xc[j * XS - T::get(j)]);
    if (WRAP) {
#pragma unroll
      for (int j = W - 1; j >= 0; --j)  // wrapped edges: rows c + N - t_j
        if (cmin + N - T::get(j) <= KMAX - 1) acc = __fadd_rn(acc, xc[j * XS + N - T::get(j)]);
    }
    bool neg = false;
    if (c < N) {
      sbuf[c] = acc;
      if (WRAP && c < S::TMAX) sbuf[c + N] = acc;
      neg = __fadd_rn(acc, ybuf[c]) < 0.0f;
    }
    const unsigned bal = __ballot_sync(kFull, neg);
    if (lane == 0) bword[cp * S::WPF + WARP] = bal;
  }
  (void)k;
}
template <class S, int WARP>
__device__ __forceinline__ void cta_gather_dispatch(int warp, int tid, int k, const float *xs, const float *ybuf, float *sbuf, unsigned *bword) {
  if constexpr (WARP < S::WPF) {
    if (warp == WARP) cta_gather_columns<S, WARP>(tid, k, xs, ybuf, sbuf, bword);
    else cta_gather_dispatch<S, WARP + 1>(warp, tid, k, xs, ybuf, sbuf, bword);
  }
}

template <class S, int VN>
__global__ void __launch_bounds__(S::THREADS, ms_cta_min_blocks<S, VN>()) ms_cyclic_cta_kernel(const __grid_constant__ MsParams p) {
  constexpr int N = S::N, W = S::W, RPL = S::RPL, NPW = S::NPW, THREADS = S::THREADS, CPASS = S::CPASS;
  constexpr bool WRAP = S::WRAP, SC = VN == VN_SC;
  constexpr int NPAD = NPW * 32;
  constexpr bool GATHER = S::GATHER;
  constexpr bool DUP = GATHER && WRAP;     // y / S mirrored behind entry N - 1: no wrap test in the row phase
  constexpr int XS = S::XS, YW = S::YW;
  constexpr int YWANT = W <= CCGPU_MS_CTA_SMALL_W ? CCGPU_MS_CTA_YN : CCGPU_MS_CTA_BIG_YN;
  constexpr int YN = (YWANT > 0 && GATHER && RPL == 1 && VN != VN_SC) ? (YWANT < W ? YWANT : W) : 0;
  using T = typename S::taps;
  extern __shared__ float xdyn[];  // GATHER: 32 guard floats, then W arrays of XS floats (rows, then a zero guard band)
  float *const xs = xdyn + 32;
  __shared__ float ys[2 * YW];  // y[YW] then S[YW]: edge (row, tap) is yrow[tap] / yrow[YW + tap]
  float *const ybuf = ys;
  float *const sbuf = ys + YW;
  __shared__ unsigned bword[CPASS * S::WPF];
  __shared__ long long s_frame;
  constexpr int NBLK = (N + 3) >> 2;
  __shared__ __align__(16) float stage[S::WPF][4 * NBLK];  // grouped mode: the channel values every warp generated
  __shared__ int s_hard[S::WPF];                            // ... and whether its frame needs the decoder
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = S::K > 0 ? S::K : p.k;

  int row[RPL];
  bool rvalid[RPL];
  float *yrow[RPL];
  unsigned rmask[RPL][NPW];
#pragma unroll
  for (int i = 0; i < RPL; ++i) {
    row[i] = tid + THREADS * i;
    rvalid[i] = row[i] < k;
    // threads without a row compute on the addresses of a row that keeps them on their own shared-memory bank
    // (row 0 would make them collide with the warp's lane 0 on every load) and store nothing
    if (!rvalid[i]) row[i] = k >= 32 ? lane : 0;
    yrow[i] = ybuf + row[i];
#pragma unroll
    for (int w = 0; w < NPW; ++w) rmask[i][w] = 0;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      int c = row[i] + T::get(j);
      if (WRAP && c >= N) c -= N;
#pragma unroll
      for (int w = 0; w < NPW; ++w)
        if ((c >> 5) == w && rvalid[i]) rmask[i][w] |= 1u << (c & 31);
    }
  }

  float r[RPL][W];
  float qold[SC ? RPL : 1][SC ? W : 1];
  if (GATHER) {
    for (int x = tid; x < 32 + W * XS; x += THREADS) xdyn[x] = 0.0f;  // what no row writes must read +0.0 for good
  }
  float *const xrow = xs + tid;  // GATHER: this thread's slot in every tap's array
  auto put_y = [&](int c, float v) {
    ybuf[c] = v;
    if (DUP && c < S::TMAX) ybuf[c + N] = v;
  };
  unsigned long long cnt[6] = { 0, 0, 0, 0, 0, 0 };
  long long frame = blockIdx.x;
  // grouped mode (see the header): the queue hands out groups of WPF frames, `frame` is the group index until a frame
  // of the group is picked for decoding
  const bool grouped = p.quick_hint != 0 && p.src == SRC_PHILOX && p.counters != nullptr && p.bits == nullptr && p.L == nullptr && p.iter == nullptr &&
                       p.failed == nullptr && p.packed == nullptr && p.status == nullptr && p.stop_rule != STOP_NONE && p.max_iter >= 1;
  long long grp = blockIdx.x;
  int w_next = S::WPF;      // next staging row to look at; WPF: screen a new group first
  bool first_group = true;
  unsigned quick_frames = 0;  // per warp: frames of this warp decided by the screening

  while (grouped || frame < static_cast<long long>(p.frames)) {
    // ---------------- frame source
    if (grouped) {
      bool found = false;
      for (;;) {
        if (w_next >= S::WPF) {
          if (!first_group) {
            if (tid == 0) s_frame = static_cast<long long>(gridDim.x) + static_cast<long long>(atomicAdd(p.work, 1ull));
            __syncthreads();
            grp = s_frame;
          }
          first_group = false;
          if (grp * S::WPF >= static_cast<long long>(p.frames)) break;  // block-uniform
          const long long fw = grp * S::WPF + warp;
          bool hard = false;
          if (fw < static_cast<long long>(p.frames)) {
            bool pos = true;
#pragma unroll
            for (int b0 = 0; b0 < NBLK; b0 += 32) {
              const int b = b0 + lane;
              if (b < NBLK) {
                const float4 v = awgn_block(p.keys, p.point, p.frame0 + static_cast<uint64_t>(fw), b, p.sigma);
                const float4 sv = make_float4(v.x * p.llr_scale, v.y * p.llr_scale, v.z * p.llr_scale, v.w * p.llr_scale);
                pos &= sv.x > 0.0f && (4 * b + 1 >= N || sv.y > 0.0f) && (4 * b + 2 >= N || sv.z > 0.0f) && (4 * b + 3 >= N || sv.w > 0.0f);
                *reinterpret_cast<float4 *>(&stage[warp][4 * b]) = sv;
              }
            }
            hard = !__all_sync(kFull, pos);
            if (!hard) ++quick_frames;
          }
          if (lane == 0) s_hard[warp] = hard ? 1 : 0;
          __syncthreads();
          w_next = 0;
        }
        while (w_next < S::WPF && !s_hard[w_next]) ++w_next;
        if (w_next < S::WPF) {
          found = true;
          break;
        }
      }
      if (!found) break;
      frame = grp * S::WPF + w_next;
      for (int c = tid; c < N; c += THREADS) put_y(c, stage[w_next][c]);
      ++w_next;
    } else if (p.src == SRC_HBM) {
      for (int c = tid; c < N; c += THREADS) put_y(c, __ldg(p.y + frame * N + c));
    } else if (p.src == SRC_PHILOX) {
      for (int b = tid; b < NBLK; b += THREADS) {
        const float4 v = awgn_block(p.keys, p.point, p.frame0 + static_cast<uint64_t>(frame), b, p.sigma);
        const float vv[4] = { v.x * p.llr_scale, v.y * p.llr_scale, v.z * p.llr_scale, v.w * p.llr_scale };
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (4 * b + e < N) put_y(4 * b + e, vv[e]);
      }
    } else if (tid == 0) {
      unsigned long long rank = p.frame0 + static_cast<unsigned long long>(frame);
      unsigned ones = p.flip_weight;
      for (int c = 0; c < N; ++c) {
        const unsigned long long zero_first = binom(N - c - 1, ones);
        float v = 1.0f;
        if (rank >= zero_first && ones > 0) {
          rank -= zero_first;
          --ones;
          v = -1.0f;
        }
        put_y(c, v);
      }
    }
    for (int c = tid; c < YW; c += THREADS) sbuf[c] = 0.0f;
#pragma unroll
    for (int i = 0; i < RPL; ++i)
#pragma unroll
      for (int j = 0; j < W; ++j) {
        r[i][j] = 0.0f;
        if (SC) qold[i][j] = 0.0f;
      }
    __syncthreads();
    // all channel values positive: iteration 0 decides the all-zero word and stops (see quick_ok in ms_cyclic.cuh);
    // nothing has to be executed for it unless the totals L are wanted
    bool quick = false;
    if (!grouped) {  // grouped mode: the screening already knows that this frame has a non-positive value
      bool allpos = true;
      for (int c = tid; c < N; c += THREADS) allpos &= ybuf[c] > 0.0f;
      quick = __syncthreads_and(allpos) && p.L == nullptr && p.stop_rule != STOP_NONE && p.max_iter >= 1;
    }
    if (quick) {  // block-uniform
      if (tid < CPASS * S::WPF) bword[tid] = 0u;
      __syncthreads();
    }

    // y of the first YN taps of this thread's row is loop invariant: keep it in registers (the shapes with few messages
    // per thread have the registers to spare; one shared-memory load less per such edge and iteration)
    float yreg[YN > 0 ? YN : 1];
    if (YN > 0 && !quick) {
#pragma unroll
      for (int j = 0; j < YN; ++j) yreg[j] = yrow[0][T::get(j)];
    }

    int it = 0;
    bool stop = quick;
    for (; !quick; ++it) {
      // ============ VN + CN
      float f1s[RPL], f2s[RPL], m1v[RPL];
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        float m1 = FLT_MAX, m2 = FLT_MAX;
        unsigned par = 0;
#pragma unroll
        for (int j = 0; j < W; ++j) {
          int off = T::get(j);
          if (WRAP && !DUP && row[i] + off >= N) off -= N;
          const float s = yrow[i][YW + off];
          const float yy = (YN > 0 && j < YN) ? yreg[j < YN ? j : 0] : yrow[i][off];
          float e = __fsub_rn(s, r[i][j]);  // scalar adds here: the packed FADD2 form of ms_cyclic.cuh needs aligned
          if (VN == VN_2D) e = __fmul_rn(p.beta_f, e);  // register pairs and spills at this kernel's 128-register budget
          float q = __fadd_rn(e, yy);
          if (SC) {
            const float qo = qold[i][j];
            if (p.variant == V_SCMS1) {
              const bool keep = (qo == 0.0f) || ((qo > 0.0f) == (q > 0.0f) && (qo < 0.0f) == (q < 0.0f));
              q = keep ? q : 0.0f;
            } else {
              q = (__fmul_rn(q, qo) > 0.0f) ? q : __fmul_rn(0.5f, __fadd_rn(q, qo));
            }
            qold[i][j] = q;
          } else {
            r[i][j] = q;
          }
          const float a = fabsf(q);
          m2 = fminf(m2, fmaxf(m1, a));
          m1 = fminf(m1, a);
          par ^= __float_as_uint(q);
        }
        m1v[i] = m1;
        float2 g = cn_magnitude_pair(p, m1, m2);
        if (GATHER && !rvalid[i]) g = make_float2(0.0f, 0.0f);  // no row: its parked messages are +-0.0, which no sum sees
        f1s[i] = xor_sign(g.x, par);
        f2s[i] = xor_sign(g.y, par);
      }
#pragma unroll
      for (int i = 0; i < RPL; ++i)
#pragma unroll
        for (int j = 0; j < W; ++j) {
          const float q = SC ? qold[i][j] : r[i][j];
          const float f = (fabsf(q) == m1v[i]) ? f2s[i] : f1s[i];
          r[i][j] = xor_sign(f, __float_as_uint(q));
          if (GATHER) xrow[j * XS] = r[i][j];  // unconditional: a thread without a row parks +-0.0 (see f1s above)
        }
      __syncthreads();
      // ============ column sums, rows ascending
      if (GATHER) {
        cta_gather_dispatch<S, 0>(warp, tid, k, xs, ybuf, sbuf, bword);
      } else {
      for (int c = tid; c < YW; c += THREADS) sbuf[c] = 0.0f;
      __syncthreads();
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
            if (rvalid[i] && wrapped == (pass == 1)) yrow[i][YW + off] = __fadd_rn(yrow[i][YW + off], r[i][j]);
          }
          __syncthreads();
        }
      }
      // ============ totals, hard decision, stop test
#pragma unroll
      for (int cp = 0; cp < CPASS; ++cp) {
        const int c = cp * THREADS + tid;
        bool neg = false;
        if (c < N) neg = __fadd_rn(sbuf[c], ybuf[c]) < 0.0f;
        const unsigned bal = __ballot_sync(kFull, neg);
        if (lane == 0) bword[cp * S::WPF + warp] = bal;
      }
      }  // !GATHER
      __syncthreads();
      bool bad = false;
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        int ov = 0;
#pragma unroll
        for (int w = 0; w < NPW; ++w) ov += __popc(bword[w] & rmask[i][w]);
        if (rvalid[i]) {
          if (p.stop_rule == STOP_REF) bad |= (ov & 255) != 0;
          else if (p.stop_rule == STOP_GF2) bad |= (ov & 1) != 0;
          else bad = true;
        }
      }
      stop = __syncthreads_or(bad) == 0;
      if (stop || it + 1 >= p.max_iter) break;
    }

    // ---------------- outputs of this frame
    const bool failed = !stop && p.stop_rule != STOP_NONE;
    int nbits = 0;
#pragma unroll
    for (int w = 0; w < NPW; ++w) nbits += __popc(bword[w]);
    if (p.bits)
      for (int c = tid; c < N; c += THREADS) p.bits[frame * N + c] = static_cast<uint8_t>((bword[c >> 5] >> (c & 31)) & 1u);
    if (p.L)
      for (int c = tid; c < N; c += THREADS) p.L[frame * N + c] = __fadd_rn(sbuf[c], ybuf[c]);
    if (p.packed && tid < NPW) p.packed[frame * NPW + tid] = bword[tid];  // columns >= N never vote: the tail bits are 0
    if (tid == 0) {
      if (p.iter) p.iter[frame] = static_cast<uint8_t>(failed ? p.max_iter : it);
      if (p.failed) p.failed[frame] = failed ? 1 : 0;
      if (p.status) p.status[frame] = static_cast<uint8_t>(failed ? 255 : it);
      cnt[C_FRAMES] += 1;
      cnt[C_ITER] += static_cast<unsigned>(it + 1);
      cnt[C_FAIL] += failed ? 1 : 0;
      cnt[C_BIT_ERR] += static_cast<unsigned>(nbits);
      cnt[C_FRAME_ERR] += (failed || nbits != 0) ? 1 : 0;
      cnt[C_UNDETECTED] += (!failed && nbits != 0) ? 1 : 0;
      if (!grouped) s_frame = static_cast<long long>(gridDim.x) + static_cast<long long>(atomicAdd(p.work, 1ull));
    }
    __syncthreads();
    if (!grouped) {
      frame = s_frame;
      __syncthreads();
    }
  }
  if (tid == 0 && p.counters != nullptr)
    for (int s = 0; s < 6; ++s)
      if (cnt[s]) atomicAdd(p.counters + s, cnt[s]);
  if (grouped && lane == 0 && quick_frames) {  // iteration 0 decided them: one iteration each, no errors
    atomicAdd(p.counters + C_FRAMES, static_cast<unsigned long long>(quick_frames));
    atomicAdd(p.counters + C_ITER, static_cast<unsigned long long>(quick_frames));
  }
}

}  // namespace ccgpu
