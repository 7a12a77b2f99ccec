// ms_cyclic_cta_q.cuh -- K2cq: the fixed-point min-sum of ms_cyclic_q.cuh for codes whose rows do not fit one warp
// (BCH(255,131): 124 rows of weight 68): one CTA of WPF warps owns TWO frames (the 16-bit halves of every word),
// thread <-> parity-check row(s).
//
// Same integer arithmetic and the same outputs as ms_cyclic_q_kernel (restatement: oracle_min_sum_fixed in
// oracle/ms_oracle.c; structure of the reference's codes/soft_decision.h:161-202).  What integer arithmetic buys
// here on top of the two frames per thread: the column sum needs no order, so every warp adds the messages of ITS
// rows into a warp-private partial accumulator (a warp-synchronous read-modify-write chain, no block barrier),
// and one pass adds the WPF partials to y.  The float kernel (ms_cyclic_cta.cuh) must keep the reference's
// summation order and pays one __syncthreads per tap (68 per iteration) for it; this one pays four (six with the general stop test).
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "channel.cuh"
#include "ms_cyclic_q.cuh"
#include "ms_params.h"
#include "ms_shape_cta.h"

namespace ccgpu {

// packed messages per thread -> resident CTAs per SM the register allocation aims at
#ifndef CCGPU_CTA_Q_SMALL_MINBLK
#define CCGPU_CTA_Q_SMALL_MINBLK 6  /* rows of weight <= 32 (the 127-row BCH(127,64)): resident CTAs per SM asked for. \
   Measured NMS_Q 5 dB: 8 (64 registers, 440 bytes spilled) 2.00e7 frames/s, 7 (72) 2.16e7, 6 (80) 2.24e7 */
#endif
template <class S> constexpr int ms_cta_q_min_blocks() { return S::RPL * S::W <= 32 ? CCGPU_CTA_Q_SMALL_MINBLK : S::RPL * S::W <= 72 ? 4 : 2; }

template <class S>
__global__ void __launch_bounds__(S::THREADS, ms_cta_q_min_blocks<S>()) ms_cyclic_cta_q_kernel(const __grid_constant__ MsParams p) {
  constexpr int N = S::N, W = S::W, RPL = S::RPL, NPW = S::NPW, THREADS = S::THREADS, CPASS = S::CPASS, WPF = S::WPF;
  constexpr bool WRAP = S::WRAP;
  constexpr int NPAD = NPW * 32;
  constexpr unsigned SIGN2 = 0x80008000u;
  using T = typename S::taps;
  __shared__ unsigned ybuf[NPAD];        // quantised channel values of both slots (fp16x2)
  __shared__ unsigned sbuf[NPAD];        // S' = y + sum_rows r
  __shared__ unsigned part[WPF][NPAD];   // per-warp partial column sums
  __shared__ unsigned bword[2][CPASS * WPF];
  __shared__ long long s_next[2];
  constexpr int NBLK = (N + 3) >> 2;
  // grouped mode (as in ms_cyclic_cta.cuh): frames are drawn in groups of WPF, one per warp; the quantised channel values
  // are parked here, all-positive frames are counted by their warp and never reach the decoder
  __shared__ __align__(8) unsigned short stage16[WPF][4 * NBLK];
  __shared__ int s_hard[WPF];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = S::K > 0 ? S::K : p.k;
  unsigned short *const ybuf16 = reinterpret_cast<unsigned short *>(ybuf);
  unsigned short *const sbuf16 = reinterpret_cast<unsigned short *>(sbuf);

  int row[RPL];
  bool rvalid[RPL];
  unsigned rmask[RPL][NPW];
#pragma unroll
  for (int i = 0; i < RPL; ++i) {
    row[i] = tid + THREADS * i;
    rvalid[i] = row[i] < k;
    if (!rvalid[i]) row[i] = 0;
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

  unsigned r[RPL][W];
#pragma unroll
  for (int i = 0; i < RPL; ++i)
#pragma unroll
    for (int j = 0; j < W; ++j) r[i][j] = 0u;
  unsigned long long cnt[6] = { 0, 0, 0, 0, 0, 0 };
  // block-uniform slot state, replicated in every thread
  long long frame[2];
  bool active[2], need_init[2];
  int it[2] = { 0, 0 };
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    frame[h] = 2 * static_cast<long long>(blockIdx.x) + h;
    active[h] = frame[h] < static_cast<long long>(p.frames);
    need_init[h] = true;
  }
  const long long units = 2 * static_cast<long long>(gridDim.x);
  for (int c = tid; c < NPAD; c += THREADS) {
    ybuf[c] = 0u;
    sbuf[c] = 0u;
  }
  const unsigned kMmax = p.q_h2_mmax, kAlpha = p.q_h2_alpha, k1024 = p.q_h2_1024, k1024B = p.q_h2_1024b;  // fp16x2, host-made
  const bool quick_ok = p.quick_hint != 0 && p.L == nullptr && p.stop_rule != STOP_NONE && p.max_iter >= 1;
  const bool grouped = quick_ok && p.src == SRC_PHILOX && p.counters != nullptr && p.bits == nullptr && p.iter == nullptr &&
                       p.failed == nullptr && p.packed == nullptr && p.status == nullptr;
  long long grp = blockIdx.x;  // grouped mode: the queue hands out groups of WPF frames
  int w_next = WPF;            // next staging row to look at; WPF: screen a new group first
  bool first_group = true, exhausted = false;
  unsigned quick_frames = 0;   // per warp: frames of this warp decided by the screening
  if (grouped) active[0] = active[1] = static_cast<long long>(blockIdx.x) * WPF < static_cast<long long>(p.frames);
  __syncthreads();

  while (active[0] || active[1]) {
    // ---------------- (re)fill; fresh frames whose quantised channel values are all positive are retired on the spot
    // (decided by iteration 0 without executing it, see ms_cyclic_q.cuh)
    while (true) {
      const bool need[2] = { active[0] && need_init[0], active[1] && need_init[1] };  // block-uniform
      if (!need[0] && !need[1]) break;
      if (grouped) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (!need[h]) continue;
          long long fr = -1;  // the next frame that needs the decoder (block-uniform search)
          while (!exhausted) {
            if (w_next >= WPF) {
              if (!first_group) {
                if (tid == 0) s_next[0] = static_cast<long long>(gridDim.x) + static_cast<long long>(atomicAdd(p.work, 1ull));
                __syncthreads();
                grp = s_next[0];
              }
              first_group = false;
              if (grp * WPF >= static_cast<long long>(p.frames)) {
                exhausted = true;
                break;
              }
              const long long fw = grp * WPF + warp;
              bool hard = false;
              if (fw < static_cast<long long>(p.frames)) {
                bool pos = true;
#pragma unroll
                for (int b0 = 0; b0 < NBLK; b0 += 32) {
                  const int b = b0 + lane;
                  if (b < NBLK) {
                    const float4 v = awgn_block(p.keys, p.point, p.frame0 + static_cast<uint64_t>(fw), b, p.sigma);
                    const unsigned short q0 = quantise_h(v.x, p.q_scale, p.q_ymax), q1 = quantise_h(v.y, p.q_scale, p.q_ymax),
                                         q2 = quantise_h(v.z, p.q_scale, p.q_ymax), q3 = quantise_h(v.w, p.q_scale, p.q_ymax);
                    pos &= static_cast<short>(q0) > 0 && (4 * b + 1 >= N || static_cast<short>(q1) > 0) &&
                           (4 * b + 2 >= N || static_cast<short>(q2) > 0) && (4 * b + 3 >= N || static_cast<short>(q3) > 0);
                    *reinterpret_cast<uint2 *>(&stage16[warp][4 * b]) = make_uint2(q0 | (static_cast<unsigned>(q1) << 16), q2 | (static_cast<unsigned>(q3) << 16));
                  }
                }
                hard = !__all_sync(kFull, pos);
                if (!hard) ++quick_frames;
              }
              if (lane == 0) s_hard[warp] = hard ? 1 : 0;
              __syncthreads();
              w_next = 0;
            }
            while (w_next < WPF && !s_hard[w_next]) ++w_next;
            if (w_next < WPF) {
              fr = grp * WPF + w_next;
              break;
            }
          }
          if (fr < 0) {
            active[h] = false;
            continue;
          }
          frame[h] = fr;
          for (int c = tid; c < N; c += THREADS) {
            const unsigned short v = stage16[w_next][c];
            ybuf16[2 * c + h] = v;
            sbuf16[2 * c + h] = v;
          }
          ++w_next;
          const unsigned keep = h == 0 ? 0xffff0000u : 0x0000ffffu;
#pragma unroll
          for (int i = 0; i < RPL; ++i)
#pragma unroll
            for (int j = 0; j < W; ++j) r[i][j] &= keep;
          it[h] = 0;
          need_init[h] = false;
        }
        __syncthreads();
        break;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (!need[h]) continue;
        if (p.src == SRC_HBM) {
          for (int c = tid; c < N; c += THREADS) {
            const unsigned short v = quantise_h(__ldg(p.y + frame[h] * N + c), p.q_scale, p.q_ymax);
            ybuf16[2 * c + h] = v;
            sbuf16[2 * c + h] = v;
          }
        } else if (p.src == SRC_PHILOX) {
          for (int b = tid; b < NBLK; b += THREADS) {
            const float4 v = awgn_block(p.keys, p.point, p.frame0 + static_cast<uint64_t>(frame[h]), b, p.sigma);
            const float vv[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (4 * b + e < N) {
                const unsigned short q = quantise_h(vv[e], p.q_scale, p.q_ymax);
                ybuf16[2 * (4 * b + e) + h] = q;
                sbuf16[2 * (4 * b + e) + h] = q;
              }
          }
        } else if (tid == 0) {
          unsigned long long rank = p.frame0 + static_cast<unsigned long long>(frame[h]);
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
            ybuf16[2 * c + h] = v;
            sbuf16[2 * c + h] = v;
          }
        }
        const unsigned keep = h == 0 ? 0xffff0000u : 0x0000ffffu;
#pragma unroll
        for (int i = 0; i < RPL; ++i)
#pragma unroll
          for (int j = 0; j < W; ++j) r[i][j] &= keep;
        it[h] = 0;
        need_init[h] = false;
      }
      __syncthreads();
      if (!quick_ok) break;
      bool again = false;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (!need[h]) continue;
        bool allpos = true;
        for (int c = tid; c < N; c += THREADS) allpos &= static_cast<short>(ybuf16[2 * c + h]) > 0;
        if (!__syncthreads_and(allpos)) continue;  // block-uniform
        if (p.bits)
          for (int c = tid; c < N; c += THREADS) p.bits[frame[h] * N + c] = 0;
        if (p.packed && tid < NPW) p.packed[frame[h] * NPW + tid] = 0u;
        if (tid == 0) {
          if (p.iter) p.iter[frame[h]] = 0;
          if (p.failed) p.failed[frame[h]] = 0;
          if (p.status) p.status[frame[h]] = 0;
          cnt[C_FRAMES] += 1;
          cnt[C_ITER] += 1;
          s_next[h] = units + static_cast<long long>(atomicAdd(p.work, 1ull));
        }
        __syncthreads();
        frame[h] = s_next[h];
        active[h] = frame[h] < static_cast<long long>(p.frames);
        need_init[h] = true;
        again = true;
      }
      if (!again) break;
      __syncthreads();  // s_next is rewritten by the next round
    }
    if (!active[0] && !active[1]) break;

    // ---------------- VN + CN, both slots (see ms_cyclic_q.cuh)
    unsigned f2s[RPL], ds[RPL], m1v[RPL];
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      unsigned m1 = 0x7bff7bffu, m2 = 0x7bff7bffu, par = 0;
#pragma unroll
      for (int j = 0; j < W; ++j) {
        int off = T::get(j);
        if (WRAP && row[i] + off >= N) off -= N;
        const unsigned q = h2_sub(sbuf[row[i] + off], r[i][j]);
        r[i][j] = q;
        const unsigned a = h2_abs(q);
        m2 = h2_min(m2, h2_max(m1, a));
        m1 = h2_min(m1, a);
        par ^= q;
      }
      const unsigned g1 = h2_max(h2_sub(h2_fma(h2_min(m1, kMmax), kAlpha, k1024), k1024B), 0u);
      const unsigned g2 = h2_max(h2_sub(h2_fma(h2_min(m2, kMmax), kAlpha, k1024), k1024B), 0u);
      const unsigned f1 = g1 ^ (par & SIGN2), f2 = g2 ^ (par & SIGN2);
      f2s[i] = f2;
      ds[i] = h2_sub(f1, f2);
      m1v[i] = m1;
    }
#pragma unroll
    for (int i = 0; i < RPL; ++i)
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const unsigned q = r[i][j];
        r[i][j] = h2_fma(ds[i], h2_gt_abs(q, m1v[i]), f2s[i]) ^ (q & SIGN2);
      }

    // ---------------- column sums: warp-private partials, then one reduction over the warps
    unsigned *const mine = part[warp];
#pragma unroll
    for (int c = lane; c < NPAD; c += 32) mine[c] = 0u;
    __syncwarp();
#pragma unroll
    for (int j = W - 1; j >= 0; --j) {
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        int off = T::get(j);
        if (WRAP && row[i] + off >= N) off -= N;
        if (rvalid[i]) {
          if (RPL == 1) {  // a warp's 32 rows hit 32 distinct columns per step; program order through volatile accesses
            volatile unsigned *sp = mine + row[i] + off;
            *sp = h2_add(*sp, r[i][j]);
          } else {
            mine[row[i] + off] = h2_add(mine[row[i] + off], r[i][j]);
          }
        }
        if (RPL > 1) __syncwarp();  // rows THREADS apart may meet in one column (wrap-around): keep the steps apart
      }
    }
    __syncthreads();  // every VN read of S' and every partial is done
#pragma unroll
    for (int cp = 0; cp < CPASS; ++cp) {
      const int c = cp * THREADS + tid;
      unsigned x = 0u;
      if (c < N) {
        x = ybuf[c];
#pragma unroll
        for (int w = 0; w < WPF; ++w) x = h2_add(x, part[w][c]);
        sbuf[c] = x;
      }
      const unsigned b0 = __ballot_sync(kFull, (x & 0x8000u) != 0u), b1 = __ballot_sync(kFull, (x & 0x80000000u) != 0u);
      if (lane == 0) {
        bword[0][cp * WPF + warp] = b0;
        bword[1][cp * WPF + warp] = b1;
      }
    }
    __syncthreads();

    // ---------------- stop test per slot
    unsigned badbits = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      bool bad = false;
      if (p.stop_simple) {
        unsigned anyone = 0;
#pragma unroll
        for (int w = 0; w < NPW; ++w) anyone |= bword[h][w];
        bad = anyone != 0u;  // block-uniform already
      } else {
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
          int ov = 0;
#pragma unroll
          for (int w = 0; w < NPW; ++w) ov += __popc(bword[h][w] & rmask[i][w]);
          if (rvalid[i]) {
            if (p.stop_rule == STOP_REF) bad |= (ov & 255) != 0;
            else if (p.stop_rule == STOP_GF2) bad |= (ov & 1) != 0;
            else bad = true;
          }
        }
      }
      badbits |= bad ? (1u << h) : 0u;
    }
    if (!p.stop_simple) {  // __syncthreads_or returns a truth value, not the bitwise OR: one reduction per slot
      const unsigned b0 = __syncthreads_or(static_cast<int>(badbits & 1u)) ? 1u : 0u;
      const unsigned b1 = __syncthreads_or(static_cast<int>(badbits & 2u)) ? 2u : 0u;
      badbits = b0 | b1;
    }

    // ---------------- outputs of the finishing slots, next frames
    bool fin[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const bool stop = ((badbits >> h) & 1u) == 0u;
      fin[h] = active[h] && (stop || it[h] + 1 >= p.max_iter);
      if (fin[h]) {
        const bool failed = !stop && p.stop_rule != STOP_NONE;
        if (p.bits)
          for (int c = tid; c < N; c += THREADS) p.bits[frame[h] * N + c] = static_cast<uint8_t>((bword[h][c >> 5] >> (c & 31)) & 1u);
        if (p.L)
          for (int c = tid; c < N; c += THREADS) p.L[frame[h] * N + c] = __half2float(__ushort_as_half(sbuf16[2 * c + h]));
        if (p.packed && tid < NPW) p.packed[frame[h] * NPW + tid] = bword[h][tid];
        if (tid == 0) {
          if (p.status) p.status[frame[h]] = static_cast<uint8_t>(failed ? 255 : it[h]);
          int nbits = 0;
#pragma unroll
          for (int w = 0; w < NPW; ++w) nbits += __popc(bword[h][w]);
          if (p.iter) p.iter[frame[h]] = static_cast<uint8_t>(failed ? p.max_iter : it[h]);
          if (p.failed) p.failed[frame[h]] = failed ? 1 : 0;
          cnt[C_FRAMES] += 1;
          cnt[C_ITER] += static_cast<unsigned>(it[h] + 1);
          cnt[C_FAIL] += failed ? 1 : 0;
          cnt[C_BIT_ERR] += static_cast<unsigned>(nbits);
          cnt[C_FRAME_ERR] += (failed || nbits != 0) ? 1 : 0;
          cnt[C_UNDETECTED] += (!failed && nbits != 0) ? 1 : 0;
          if (!grouped) s_next[h] = units + static_cast<long long>(atomicAdd(p.work, 1ull));
        }
      }
    }
    if (fin[0] || fin[1]) {  // block-uniform
      __syncthreads();
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (fin[h]) {
          if (!grouped) {
            frame[h] = s_next[h];
            active[h] = frame[h] < static_cast<long long>(p.frames);
          }
          need_init[h] = true;  // grouped mode: the refill finds the slot's next frame (or ends the slot)
        }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (!fin[h]) ++it[h];
    // the refill (or the next VN pass) rewrites / reads what this iteration's tail has read: one barrier covers both
    __syncthreads();
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
