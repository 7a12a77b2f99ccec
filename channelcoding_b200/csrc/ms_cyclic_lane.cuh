// ms_cyclic_lane.cuh -- K2s: flooding min-sum / sum-product for SMALL cyclic parity-check matrices (n = 15, n = 31),
// sm_100a: one LANE owns one frame.
//
// Same arithmetic, same order, same outputs as ms_cyclic_kernel (reference codes/soft_decision.h:161-202:
// vertical__ :125-140, horizontal__ :101-122, column_sum :86-98 rows ascending, syndrome :79-84); what changes is the
// mapping.  With lane <-> row (ms_cyclic.cuh) a BCH(15,7) frame occupies 8 lanes for 4 edges each, and an iteration of
// 32 edges pays the same ballots, barriers, queue bookkeeping and shared-memory round trips as one of 486: 238 warp
// instructions per frame at 3 dB, of which the edges are a quarter.  Here the whole decoder state of a frame -- K x W
// messages, y[n], the column sums S[n] -- lives in the registers of ONE lane, every loop is unrolled over rows and taps,
// the column of edge (row, tap) is a compile-time constant, so there is no shared-memory traffic, no shuffle and no
// vote inside an iteration: a warp instruction advances 32 frames at once.
//   * frames finish after different iteration counts: a lane that finishes takes the next frame from its warp's FIFO in
//     shared memory at once (the other lanes keep iterating), so no lane waits for the slowest frame of a batch.
//   * the FIFO is filled by the whole warp: Philox blocks, one per lane (8 frames of n = 15 per pass: all lanes busy), or
//     coalesced loads of up to 32 consecutive frames from HBM.
//   * frame indices come from the global queue head in batches (one atomic per up to 8 x work_batch frames).
// n = 31 (80 .. 120 messages + 93 column values per lane): 255 registers with 120 .. 300 bytes of spills, two CTAs per SM --
// eight warps per SM, but the iteration needs no shared memory and has 5 .. 20 independent rows to overlap: measured
// 1.3 .. 4.0 x the warp kernel for the min-sum flavours (BCH(31,26) 4 dB 8.0e8 -> 3.2e9 frames/s), 0.9 x for sum-product,
// which therefore stays on the warp kernel for n = 31 (CCGPU_MS_LANE_SPA_LIST).
// Float semantics as in ms_cyclic.cuh: explicit _rn adds / multiplies (no contraction), OMS offset in double, the column
// sums start at +0.0f and add the rows in ascending order (the unrolled row loop IS that order).
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "channel.cuh"
#include "ms_cyclic.cuh"
#include "ms_params.h"
#include "ms_shape.h"

#ifndef CCGPU_MS_LANE_MINBLK
#define CCGPU_MS_LANE_MINBLK 4  /* resident CTAs per SM asked for: 128 registers per thread */
#endif

namespace ccgpu {

constexpr int kLaneFifo = 64;  // frames a warp's FIFO holds (a power of two >= 32 + the largest producer pass)
template <class S> constexpr int ms_lane_pad() { return S::N <= 16 ? 16 : 32; }  // floats per FIFO slot
template <class S> constexpr int ms_lane_smem_bytes() {
  return (kMsThreads / 32) * (kLaneFifo * ms_lane_pad<S>() * static_cast<int>(sizeof(float)) + kLaneFifo * static_cast<int>(sizeof(long long)));
}
template <class S> constexpr int ms_lane_min_blocks() { return S::K * S::W <= 48 ? CCGPU_MS_LANE_MINBLK : 2; }

template <class S, int VN>
__global__ void __launch_bounds__(kMsThreads, ms_lane_min_blocks<S>()) ms_cyclic_lane_kernel(const __grid_constant__ MsParams p) {
  constexpr int N = S::N, K = S::K, W = S::W;
  constexpr bool SPA = VN == VN_SPA;
  constexpr int kLanePad = S::N <= 16 ? 16 : 32;  // = ms_lane_pad<S>()
  static_assert(K > 0 && !S::WRAP && N <= kLanePad, "exact small shapes only");
  static_assert(VN == VN_PLAIN || VN == VN_2D || VN == VN_SPA, "flavours of the lane kernel");
  using T = typename S::taps;
  constexpr int NBLK = (N + 3) >> 2;
  constexpr int PASS_PHILOX = 32 / NBLK;  // frames one producer pass generates
  extern __shared__ __align__(16) unsigned char lane_smem[];
  const int lane = threadIdx.x & 31;
  const int warp_in_cta = threadIdx.x >> 5;
  float *const fifo = reinterpret_cast<float *>(lane_smem) + warp_in_cta * (kLaneFifo * kLanePad);
  long long *const fifo_idx = reinterpret_cast<long long *>(lane_smem + (kMsThreads / 32) * kLaneFifo * kLanePad * sizeof(float)) +
                              warp_in_cta * kLaneFifo;
  __shared__ unsigned cnt_s[6][kMsThreads];  // per-thread statistics (32-bit: a lane sees far fewer than 2^32 / 50 frames)
#pragma unroll
  for (int s = 0; s < 6; ++s) cnt_s[s][threadIdx.x] = 0u;

  // ---- warp-uniform queue / FIFO state (registers, identical in every lane)
  long long pool_next = 0;
  int pool_left = 0;
  unsigned pool_batch = 0;
  int head = 0, count = 0;
  bool exhausted = false;

  // ---- per-lane decoder state
  float r[K][W], y[N], s[N];
  long long frame = 0;
  int it = 0;
  bool active = false;
#pragma unroll
  for (int c = 0; c < N; ++c) y[c] = s[c] = 0.0f;
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j < W; ++j) r[i][j] = 0.0f;

  const bool is_oms = p.variant == V_OMS;
  const float fn_scale = (p.variant == V_NMS || p.variant == V_NMS2D) ? p.alpha_f : 1.0f;
  const unsigned ov_mask = p.stop_rule == STOP_REF ? 255u : (p.stop_rule == STOP_GF2 ? 1u : 0u);
  const unsigned ov_none = p.stop_rule == STOP_NONE ? 1u : 0u;  // no stop rule: never "converged"

  while (true) {
    // ================= lanes without a frame take the next ones of the FIFO
    const unsigned needm = __ballot_sync(kFull, !active);
    if (needm) {
      const int want = __popc(needm);
      while (count < want && !exhausted) {  // ---- producer pass (warp-uniform)
        if (pool_left == 0) {
          long long nx = 0;
          unsigned b = 0;
          if (lane == 0) {
            const long long rem = static_cast<long long>(p.frames) - pool_next - (static_cast<long long>(pool_batch) << (p.work_shift - 2));
            const long long g = rem > 0 ? (rem >> p.work_shift) : 0;
            const long long cap = 8ll * p.work_batch;
            b = static_cast<unsigned>(g < 1 ? 1 : (g < cap ? g : cap));
            nx = static_cast<long long>(atomicAdd(p.work, static_cast<unsigned long long>(b)));
          }
          pool_next = __shfl_sync(kFull, nx, 0);
          pool_batch = __shfl_sync(kFull, b, 0);
          pool_left = static_cast<int>(pool_batch);
        }
        const int pass = p.src == SRC_PHILOX ? PASS_PHILOX : 32;
        const int take = pool_left < pass ? pool_left : pass;
        const long long base = pool_next;
        pool_next += take;
        pool_left -= take;
        const long long room = static_cast<long long>(p.frames) - base;
        const int nvalid = room <= 0 ? 0 : (room < take ? static_cast<int>(room) : take);
        if (nvalid == 0) {  // indices are handed out in ascending order: nothing is left for this warp
          exhausted = true;
          break;
        }
        const int tail = head + count;
        if (p.src == SRC_PHILOX) {
          const int f = lane / NBLK, blk = lane - f * NBLK;
          if (f < nvalid) {
            const float4 v = awgn_block(p.keys, p.point, p.frame0 + static_cast<uint64_t>(base + f), blk, p.sigma);
            *reinterpret_cast<float4 *>(&fifo[((tail + f) & (kLaneFifo - 1)) * kLanePad + 4 * blk]) =
                make_float4(v.x * p.llr_scale, v.y * p.llr_scale, v.z * p.llr_scale, v.w * p.llr_scale);
          }
        } else {  // SRC_HBM: nvalid consecutive frames = nvalid * N consecutive floats, read coalesced
          const float *src = p.y + base * N;
#pragma unroll
          for (int e0 = 0; e0 < 32 * N; e0 += 32) {
            const int e = e0 + lane;
            if (e < nvalid * N) {
              const int f = e / N, c = e - f * N;
              fifo[((tail + f) & (kLaneFifo - 1)) * kLanePad + c] = __ldg(src + e);
            }
          }
        }
        if (lane < nvalid) fifo_idx[(tail + lane) & (kLaneFifo - 1)] = base + lane;
        count += nvalid;
      }
      __syncwarp();
      const int rank = __popc(needm & ((1u << lane) - 1u));
      if (!active && rank < count) {
        const int slot = (head + rank) & (kLaneFifo - 1);
        const float4 *src = reinterpret_cast<const float4 *>(&fifo[slot * kLanePad]);
        float tmp[kLanePad];
#pragma unroll
        for (int b = 0; b < NBLK; ++b) {
          const float4 v = src[b];
          tmp[4 * b] = v.x;
          tmp[4 * b + 1] = v.y;
          tmp[4 * b + 2] = v.z;
          tmp[4 * b + 3] = v.w;
        }
#pragma unroll
        for (int c = 0; c < N; ++c) {
          y[c] = tmp[c];
          s[c] = 0.0f;
        }
        frame = fifo_idx[slot];
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
          for (int j = 0; j < W; ++j) r[i][j] = 0.0f;
        it = 0;
        active = true;
      }
      const int taken = want < count ? want : count;
      head = (head + taken) & (kLaneFifo - 1);
      count -= taken;
      __syncwarp();  // the slots just read may be rewritten by the next producer pass
    }
    if (__ballot_sync(kFull, active) == 0u) break;

    // ================= one flooding iteration of this lane's frame (lanes without a frame run along on stale registers)
    float sn[N];
#pragma unroll
    for (int c = 0; c < N; ++c) sn[c] = 0.0f;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      float q[W];
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const int c = i + T::get(j);
        float e = __fsub_rn(s[c], r[i][j]);           // vertical__ :135-136
        if (VN == VN_2D) e = __fmul_rn(p.beta_f, e);  // normalised_vertical :215-218
        q[j] = __fadd_rn(e, y[c]);
      }
      if (SPA) {
        // extension (not in the reference), the same expressions as ms_cyclic.cuh: r_j = 2 atanh(prod_{i != j} tanh(q_i / 2))
        // as prefix * suffix, clamped so that atanh stays finite
        float th[W], pre[W];
        float prod = 1.0f;
#pragma unroll
        for (int j = 0; j < W; ++j) {
          th[j] = tanhf(0.5f * q[j]);
          pre[j] = prod;
          prod *= th[j];
        }
        float suffix = 1.0f;
#pragma unroll
        for (int j = W - 1; j >= 0; --j) {
          float pr = pre[j] * suffix;
          pr = fminf(fmaxf(pr, -0.99999994f), 0.99999994f);
          r[i][j] = 2.0f * atanhf(pr);
          suffix *= th[j];
        }
      } else {
        float m1 = FLT_MAX, m2 = FLT_MAX;
        unsigned par = 0;
#pragma unroll
        for (int j = 0; j < W; ++j) {
          const float a = fabsf(q[j]);
          m2 = fminf(m2, fmaxf(m1, a));
          m1 = fminf(m1, a);
          par ^= __float_as_uint(q[j]);
        }
        // fn_h of the variant (:204-213, :245-251).  m1 / m2 are never NaN (fminf keeps the other operand), so the
        // unnormalised variants can multiply by 1.0f exactly instead of branching per row; the offset rule (double
        // intermediate) is the rare, warp-uniform case
        float2 g = make_float2(__fmul_rn(fn_scale, m1), __fmul_rn(fn_scale, m2));
        if (is_oms) g = cn_offset_pair(p.beta_d, m1, m2);
        const float f1 = xor_sign(g.x, par), f2 = xor_sign(g.y, par);
#pragma unroll
        for (int j = 0; j < W; ++j) {
          const float f = (fabsf(q[j]) == m1) ? f2 : f1;  // min over the others; ties: min2 == min1
          r[i][j] = xor_sign(f, __float_as_uint(q[j]));   // sign = product of the other signs (:114, :118)
        }
      }
#pragma unroll
      for (int j = 0; j < W; ++j) {  // column_sum :86-98 -- row i is added after rows 0 .. i-1
        const int c = i + T::get(j);
        sn[c] = __fadd_rn(sn[c], r[i][j]);
      }
    }
    unsigned word = 0;  // hard decision (:178-183)
#pragma unroll
    for (int c = 0; c < N; ++c) {
      s[c] = sn[c];
      word |= (__fadd_rn(sn[c], y[c]) < 0.0f) ? (1u << c) : 0u;
    }
    bool stop;
    if (p.stop_simple) {
      stop = word == 0u;
    } else {
      // integer overlap of every row with the decided word: reference rule (ov mod 256 != 0), GF(2) parity, or no rule
      unsigned bad = ov_none;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        unsigned rm = 0;
#pragma unroll
        for (int j = 0; j < W; ++j) rm |= 1u << (i + T::get(j));
        bad |= static_cast<unsigned>(__popc(word & rm)) & ov_mask;
      }
      stop = bad == 0u;
    }
    const bool fin = active && (stop || it + 1 >= p.max_iter);
    if (fin) {
      const bool failed = !stop && p.stop_rule != STOP_NONE;
      const int nbits = __popc(word);
      if (p.bits) {
#pragma unroll
        for (int c = 0; c < N; ++c) p.bits[frame * N + c] = static_cast<uint8_t>((word >> c) & 1u);
      }
      if (p.L) {
#pragma unroll
        for (int c = 0; c < N; ++c) p.L[frame * N + c] = __fadd_rn(s[c], y[c]);
      }
      if (p.iter) p.iter[frame] = static_cast<uint8_t>(failed ? p.max_iter : it);
      if (p.failed) p.failed[frame] = failed ? 1 : 0;
      if (p.packed) p.packed[frame] = word;  // n <= 32: one word per frame
      if (p.status) p.status[frame] = static_cast<uint8_t>(failed ? 255 : it);
      cnt_s[C_FRAMES][threadIdx.x] += 1u;
      cnt_s[C_ITER][threadIdx.x] += static_cast<unsigned>(it + 1);
      if (failed || nbits != 0) {
        cnt_s[C_FRAME_ERR][threadIdx.x] += 1u;
        cnt_s[C_BIT_ERR][threadIdx.x] += static_cast<unsigned>(nbits);
        cnt_s[C_FAIL][threadIdx.x] += failed ? 1u : 0u;
        cnt_s[C_UNDETECTED][threadIdx.x] += failed ? 0u : 1u;
      }
      active = false;
    } else {
      ++it;
    }
  }

  if (p.counters != nullptr) {
#pragma unroll
    for (int sl = 0; sl < 6; ++sl) {
      unsigned long long x = cnt_s[sl][threadIdx.x];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
      if (lane == 0 && x) atomicAdd(p.counters + sl, x);
    }
  }
}

}  // namespace ccgpu
