// ms_csr.cu -- K2g: min-sum / sum-product decoder for an ARBITRARY dense 0/1 parity-check matrix.
//
// Same contract as ms_cyclic.cuh (reference codes/soft_decision.h:161-202) but no structural
// assumption on H: used for H_alt() (codes/cyclic.h:361-385) and any other matrix handed in through
// ccgpu_code_from_dense whose shape has no compiled cyclic kernel (row-permuted, multiple-bases, ...),
// for every variant including the sum-product (tanh rule) extension.
//
// Mapping: one CTA per frame, persistent (grid-strided frames).  All messages of the frame stay
// in shared memory:  q/r per edge in ELL layout [slot j][row r] (conflict free for thread = row),
// y[n], S[n], decisions b[n].  One iteration is three phases separated by __syncthreads:
//   A  thread <-> row      VN (soft_decision.h:125-140) + CN (:101-122) over the row's slots
//   B  thread <-> column   S_c = sum of r over the column's edges in ascending row order (:86-98,
//                          order fixed by csc_edge), L = S + y, b = L < 0 (:178-183)
//   C  thread <-> row      stop test (:79-84): integer overlap of the row with b
// Index tables (ELL columns, CSC pointers/edges; uint16) are staged in shared memory once per CTA.
#include <cfloat>
#include <vector>

#include "channel.cuh"
#include "ms_csr.h"

namespace ccgpu {

namespace {

struct CsrView {
  int rows, n, wmax, edges;
  const uint16_t *ell_col, *csc_ptr, *csc_edge;
};

__device__ __forceinline__ float xor_sign_f(float v, unsigned signbits) {
  return __uint_as_float(__float_as_uint(v) ^ (signbits & 0x80000000u));
}
__device__ __forceinline__ float cn_mag(const MsParams &p, float m) {
  if (p.variant == V_NMS || p.variant == V_NMS2D) return __fmul_rn(p.alpha_f, m);
  if (p.variant == V_OMS) {
    const double d = static_cast<double>(m) - p.beta_d;
    return static_cast<float>(d > 0.0 ? d : 0.0);
  }
  return m;
}
__device__ __forceinline__ unsigned long long binom_u64(unsigned n, unsigned r) {
  if (r > n) return 0ull;
  unsigned long long v = 1ull;
  for (unsigned i = 1; i <= r; ++i) v = v * (n - r + i) / i;
  return v;
}

__global__ void ms_csr_kernel(const __grid_constant__ MsParams p, const CsrView h) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int rows = h.rows, n = h.n, wmax = h.wmax, slots = h.wmax * h.rows;
  float *qbuf = reinterpret_cast<float *>(smem_raw);  // [slots]  q, then r of the current iteration
  float *qprev = qbuf + slots;                        // [slots]  previous q (SCMS) / prefix products (SPA)
  float *ybuf = qprev + slots;                        // [n]
  float *sbuf = ybuf + n;                             // [n]
  uint16_t *ell = reinterpret_cast<uint16_t *>(sbuf + n);  // [slots]
  uint16_t *cptr = ell + slots;                            // [n + 1]
  uint16_t *cedge = cptr + (n + 1);                        // [edges]
  uint8_t *bbuf = reinterpret_cast<uint8_t *>(cedge + h.edges);  // [n]
  __shared__ int s_flags[4];  // 0: any row violated  1: decided-bit count

  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < slots; i += nt) ell[i] = h.ell_col[i];
  for (int i = tid; i <= n; i += nt) cptr[i] = h.csc_ptr[i];
  for (int i = tid; i < h.edges; i += nt) cedge[i] = h.csc_edge[i];
  __syncthreads();

  const bool sc = p.variant == V_SCMS1 || p.variant == V_SCMS2;
  const bool is2d = p.variant == V_NMS2D;
  const bool spa = p.variant == V_SPA;
  unsigned long long cnt[6] = { 0, 0, 0, 0, 0, 0 };

  for (unsigned long long fr = blockIdx.x; fr < p.frames; fr += gridDim.x) {
    // ---------------- frame source
    if (p.src == SRC_HBM) {
      for (int c = tid; c < n; c += nt) ybuf[c] = __ldg(p.y + fr * n + c);
    } else if (p.src == SRC_PHILOX) {
      for (int b = tid; b < ((n + 3) >> 2); b += nt) {
        const float4 v = awgn_block(p.keys, p.point, p.frame0 + fr, b, p.sigma);
        const float vv[4] = { v.x * p.llr_scale, v.y * p.llr_scale, v.z * p.llr_scale, v.w * p.llr_scale };
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (4 * b + e < n) ybuf[4 * b + e] = vv[e];
      }
    } else {
      if (tid == 0) {
        unsigned long long rank = p.frame0 + fr;
        unsigned ones = p.flip_weight;
        for (int c = 0; c < n; ++c) {
          const unsigned long long zero_first = binom_u64(n - c - 1, ones);
          float v = 1.0f;
          if (rank >= zero_first && ones > 0) {
            rank -= zero_first;
            --ones;
            v = -1.0f;
          }
          ybuf[c] = v;
        }
      }
    }
    for (int c = tid; c < n; c += nt) sbuf[c] = 0.0f;
    for (int i = tid; i < slots; i += nt) {
      qbuf[i] = 0.0f;   // r = 0 (matrix<R> r value-initialised, soft_decision.h:168)
      qprev[i] = 0.0f;  // q = 0 (:167)
    }
    __syncthreads();

    int it = 0;
    bool stop = false;
    for (; it < p.max_iter; ++it) {
      // ======== phase A: rows
      for (int r = tid; r < rows; r += nt) {
        float m1 = FLT_MAX, m2 = FLT_MAX;
        unsigned par = 0;
        float prod = 1.0f;
        for (int j = 0; j < wmax; ++j) {
          const int c = ell[j * rows + r];
          if (c == 0xffff) continue;
          float e = __fsub_rn(sbuf[c], qbuf[j * rows + r]);
          if (is2d) e = __fmul_rn(p.beta_f, e);
          float q = __fadd_rn(e, ybuf[c]);
          if (sc) {
            const float qo = qprev[j * rows + r];
            if (p.variant == V_SCMS1) {
              const bool keep = (qo == 0.0f) || ((qo > 0.0f) == (q > 0.0f) && (qo < 0.0f) == (q < 0.0f));
              q = keep ? q : 0.0f;
            } else {
              q = (__fmul_rn(q, qo) > 0.0f) ? q : __fmul_rn(0.5f, __fadd_rn(q, qo));
            }
            qprev[j * rows + r] = q;
          }
          if (spa) {
            const float th = tanhf(0.5f * q);
            qprev[j * rows + r] = prod;  // product of the slots before j
            prod *= th;
            qbuf[j * rows + r] = th;
          } else {
            qbuf[j * rows + r] = q;
            const float a = fabsf(q);
            m2 = fminf(m2, fmaxf(m1, a));
            m1 = fminf(m1, a);
            par ^= __float_as_uint(q);
          }
        }
        if (spa) {
          float suffix = 1.0f;  // product of the slots after j
          for (int j = wmax - 1; j >= 0; --j) {
            if (ell[j * rows + r] == 0xffff) continue;
            const float th = qbuf[j * rows + r];
            float pr = qprev[j * rows + r] * suffix;
            pr = fminf(fmaxf(pr, -0.99999994f), 0.99999994f);
            qbuf[j * rows + r] = 2.0f * atanhf(pr);
            suffix *= th;
          }
        } else {
          const float f1 = xor_sign_f(cn_mag(p, m1), par), f2 = xor_sign_f(cn_mag(p, m2), par);
          for (int j = 0; j < wmax; ++j) {
            if (ell[j * rows + r] == 0xffff) continue;
            const float q = qbuf[j * rows + r];
            qbuf[j * rows + r] = xor_sign_f((fabsf(q) == m1) ? f2 : f1, __float_as_uint(q));
          }
        }
      }
      if (tid == 0) {
        s_flags[0] = 0;
        s_flags[1] = 0;
      }
      __syncthreads();
      // ======== phase B: columns
      int mybits = 0;
      for (int c = tid; c < n; c += nt) {
        float s = 0.0f;
        for (int i = cptr[c]; i < cptr[c + 1]; ++i) s = __fadd_rn(s, qbuf[cedge[i]]);
        sbuf[c] = s;
        const bool one = __fadd_rn(s, ybuf[c]) < 0.0f;
        bbuf[c] = one ? 1 : 0;
        mybits += one ? 1 : 0;
      }
      if (mybits) atomicAdd(&s_flags[1], mybits);
      __syncthreads();
      // ======== phase C: stop test
      bool bad = false;
      for (int r = tid; r < rows; r += nt) {
        int ov = 0;
        for (int j = 0; j < wmax; ++j) {
          const int c = ell[j * rows + r];
          if (c != 0xffff) ov += bbuf[c];
        }
        if (p.stop_rule == STOP_REF) bad |= (ov & 255) != 0;
        else if (p.stop_rule == STOP_GF2) bad |= (ov & 1) != 0;
        else bad = true;
      }
      if (bad) s_flags[0] = 1;
      __syncthreads();
      stop = s_flags[0] == 0;
      if (stop || it + 1 >= p.max_iter) break;
      __syncthreads();  // s_flags are rewritten after phase A of the next iteration
    }
    // ---------------- outputs of this frame
    const bool failed = !stop && p.stop_rule != STOP_NONE;
    const int nbits = s_flags[1];
    if (p.bits)
      for (int c = tid; c < n; c += nt) p.bits[fr * n + c] = bbuf[c];
    if (p.packed) {  // compact layout: one thread per 32-bit word of the decided word
      const int npw = (n + 31) >> 5;
      for (int w = tid; w < npw; w += nt) {
        unsigned v = 0u;
        for (int b = 0; b < 32 && 32 * w + b < n; ++b) v |= (bbuf[32 * w + b] ? 1u : 0u) << b;
        p.packed[fr * npw + w] = v;
      }
    }
    if (p.L)
      for (int c = tid; c < n; c += nt) p.L[fr * n + c] = __fadd_rn(sbuf[c], ybuf[c]);
    if (tid == 0) {
      if (p.iter) p.iter[fr] = static_cast<uint8_t>(failed ? p.max_iter : it);
      if (p.failed) p.failed[fr] = failed ? 1 : 0;
      if (p.status) p.status[fr] = static_cast<uint8_t>(failed ? 255 : it);
      cnt[C_FRAMES] += 1;
      cnt[C_ITER] += static_cast<unsigned>(failed ? p.max_iter : it + 1);
      cnt[C_FAIL] += failed ? 1 : 0;
      cnt[C_BIT_ERR] += static_cast<unsigned>(nbits);
      cnt[C_FRAME_ERR] += (failed || nbits != 0) ? 1 : 0;
      cnt[C_UNDETECTED] += (!failed && nbits != 0) ? 1 : 0;
    }
    __syncthreads();
  }
  if (tid == 0 && p.counters != nullptr)
    for (int s = 0; s < 6; ++s)
      if (cnt[s]) atomicAdd(p.counters + s, cnt[s]);
}

size_t csr_smem_bytes(int rows, int n, int wmax, int edges) {
  const size_t slots = size_t(wmax) * rows;
  size_t b = (2 * slots + 2 * size_t(n)) * sizeof(float);
  b += (slots + size_t(n) + 1 + edges) * sizeof(uint16_t);
  b += size_t(n);
  return (b + 15) & ~size_t(15);
}

}  // namespace

int ms_csr_upload(const uint8_t *H, unsigned rows, unsigned n, MsCsrDevice *out) {
  *out = MsCsrDevice();
  if (rows >= 0xffff || n >= 0xffff) return 0;  // leave empty: launch reports unsupported
  int wmax = 0, edges = 0;
  for (unsigned r = 0; r < rows; ++r) {
    int w = 0;
    for (unsigned c = 0; c < n; ++c) w += H[size_t(r) * n + c] ? 1 : 0;
    wmax = std::max(wmax, w);
    edges += w;
  }
  const size_t slots = size_t(wmax) * rows;
  if (slots >= 0xffff || wmax == 0) return 0;
  std::vector<uint16_t> ell(slots, 0xffff), cptr(n + 1, 0), cedge(edges, 0);
  std::vector<std::vector<uint16_t>> percol(n);
  for (unsigned r = 0; r < rows; ++r) {
    int j = 0;
    for (unsigned c = 0; c < n; ++c)
      if (H[size_t(r) * n + c]) {
        ell[size_t(j) * rows + r] = static_cast<uint16_t>(c);
        percol[c].push_back(static_cast<uint16_t>(size_t(j) * rows + r));  // rows ascending by construction
        ++j;
      }
  }
  int pos = 0;
  for (unsigned c = 0; c < n; ++c) {
    cptr[c] = static_cast<uint16_t>(pos);
    for (uint16_t e : percol[c]) cedge[pos++] = e;
  }
  cptr[n] = static_cast<uint16_t>(pos);
  out->rows = static_cast<int>(rows);
  out->n = static_cast<int>(n);
  out->wmax = wmax;
  out->edges = edges;
  out->smem_bytes = csr_smem_bytes(out->rows, out->n, wmax, edges);
  out->threads = std::min(256, std::max(64, int((std::max(rows, n) + 31) / 32 * 32)));
  if (cudaMalloc(&out->ell_col, slots * sizeof(uint16_t)) != cudaSuccess ||
      cudaMalloc(&out->csc_ptr, (n + 1) * sizeof(uint16_t)) != cudaSuccess ||
      cudaMalloc(&out->csc_edge, std::max(1, edges) * sizeof(uint16_t)) != cudaSuccess)
    return -1;
  if (cudaMemcpy(out->ell_col, ell.data(), slots * sizeof(uint16_t), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(out->csc_ptr, cptr.data(), (n + 1) * sizeof(uint16_t), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(out->csc_edge, cedge.data(), edges * sizeof(uint16_t), cudaMemcpyHostToDevice) != cudaSuccess)
    return -1;
  return 0;
}

void ms_csr_free(MsCsrDevice *d) {
  if (d->ell_col) cudaFree(d->ell_col);
  if (d->csc_ptr) cudaFree(d->csc_ptr);
  if (d->csc_edge) cudaFree(d->csc_edge);
  *d = MsCsrDevice();
}

int ms_csr_launch(const MsCsrDevice &d, const MsParams &mp, int sm_count, cudaStream_t stream) {
  if (!d.ell_col || d.smem_bytes > 227 * 1024) return -3;
  if (cudaFuncSetAttribute(ms_csr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(d.smem_bytes)) !=
      cudaSuccess)
    return -1;
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ms_csr_kernel, d.threads, d.smem_bytes);
  const unsigned long long cap = static_cast<unsigned long long>(std::max(1, occ)) * sm_count;
  const unsigned grid = static_cast<unsigned>(std::min<unsigned long long>(mp.frames, cap));
  CsrView v{ d.rows, d.n, d.wmax, d.edges, d.ell_col, d.csc_ptr, d.csc_edge };
  ms_csr_kernel<<<grid, d.threads, d.smem_bytes, stream>>>(mp, v);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace ccgpu
