// gf_decode.cu -- K4: batched hard-decision (algebraic) decoding of primitive BCH / RS codes
// over GF(2^q), q <= 8, one warp per received word, log/antilog tables in shared memory.
//
// Replaces  cyclic::correct_(b, erasures = {}, hard_decision_tag)   reference codes/cyclic.h:207-252:
//   syndromes  s_j = b(alpha^(root_j))          cyclic.h:53-63 + math/polynomial.h:273-284 (Horner)
//   locator    error_locator_polynomial(...)    codes/hard_decision.h:61-196 (PGZ / BM / Euklid tags)
//   roots      zeroes(sigma) by exhaustive evaluation, #roots must equal deg sigma  cyclic.h:126-150
//   values     1 for binary BCH (codes/bch.h:80-83); for RS the reference solves a linear system
//              (codes/rs.h:41-78) -- here Forney's formula, which yields the same unique values
//   fix-up + re-syndrome check, failure if any syndrome is left   cyclic.h:237-248
// The locator is computed with Berlekamp-Massey, lane-parallel over the coefficients; the
// reference's three algorithm tags all implement the same bounded-distance decoder (its own BM has
// an out-of-bounds read), so results are compared against the Euklid tag.
//
// Work split inside the warp: lane <-> syndrome index (Horner over the n symbols, the symbol is a
// shared-memory broadcast), lane <-> locator coefficient (BM), lane <-> codeword position (root
// search), lane <-> error index (Forney).
#include <algorithm>
#include <vector>

#include "gf_decode.h"

namespace ccgpu {

namespace {

constexpr int kGfThreads = 128;
constexpr int kGfWarps = kGfThreads / 32;
constexpr unsigned kAll = 0xffffffffu;

struct GfParams {
  int q, n, t, nroots, binary, mu;
  const uint8_t *tables;
  const uint8_t *words;
  unsigned long long count;
  uint8_t *corrected, *n_errors, *failed;
};

struct Gf {
  const uint8_t *exp, *log;
  __device__ __forceinline__ unsigned mul(unsigned a, unsigned b) const { return (a && b) ? exp[log[a] + log[b]] : 0u; }
  // a * alpha^e, e < n
  __device__ __forceinline__ unsigned mul_alpha(unsigned a, unsigned e) const { return a ? exp[log[a] + e] : 0u; }
};

__device__ __forceinline__ unsigned xor_reduce(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v ^= __shfl_xor_sync(kAll, v, o);
  return v;
}

// syndromes of `word` into synd[0..nroots); returns true if any is non-zero
__device__ __forceinline__ bool syndromes(const GfParams &p, const Gf &F, const uint8_t *word, const uint8_t *rexp,
                                          uint8_t *synd, int lane) {
  bool any = false;
  for (int j0 = 0; j0 < p.nroots; j0 += 32) {
    const int j = j0 + lane;
    unsigned s = 0;
    if (j < p.nroots) {
      const unsigned e = rexp[j];
      for (int i = p.n - 1; i >= 0; --i) s = F.mul_alpha(s, e) ^ word[i];
      synd[j] = static_cast<uint8_t>(s);
    }
    any |= __any_sync(kAll, s != 0);
  }
  __syncwarp();
  return any;
}

__global__ void __launch_bounds__(kGfThreads) gf_decode_kernel(const GfParams p) {
  extern __shared__ uint8_t sm[];
  const int size = 1 << p.q, n = p.n;
  uint8_t *s_exp = sm;                 // [2*size]
  uint8_t *s_log = s_exp + 2 * size;   // [size]
  uint8_t *s_rexp = s_log + size;      // [64] root exponents
  uint8_t *per_warp = s_rexp + 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t *word = per_warp + warp * 512;  // [256]
  uint8_t *synd = word + 256;             // [64]
  uint8_t *lamv = synd + 64;              // [32] locator coefficients
  uint8_t *omega = lamv + 32;             // [32]
  uint8_t *pos = omega + 32;              // [32] error positions
  for (int i = threadIdx.x; i < 3 * size; i += blockDim.x) sm[i] = p.tables[i];
  if (threadIdx.x < 64)
    s_rexp[threadIdx.x] = static_cast<uint8_t>(threadIdx.x < p.nroots ? (p.mu + threadIdx.x) % n : 0);
  __syncthreads();
  const Gf F{ s_exp, s_log };

  const unsigned long long nwarps = static_cast<unsigned long long>(gridDim.x) * kGfWarps;
  for (unsigned long long w = static_cast<unsigned long long>(blockIdx.x) * kGfWarps + warp; w < p.count; w += nwarps) {
    const uint8_t *in = p.words + w * n;
    for (int i = lane; i < n; i += 32) word[i] = in[i];
    __syncwarp();
    bool failed = false;
    int nerr = 0;
    if (syndromes(p, F, word, s_rexp, synd, lane)) {
      // ---------------- Berlekamp-Massey, coefficient `lane` of Lambda and B
      unsigned lam = lane == 0, B = lane == 0, b = 1;
      int L = 0, m = 1;
      for (int r = 0; r < p.nroots; ++r) {
        const unsigned term = (lane <= L && lane <= r) ? F.mul(lam, synd[r - lane]) : 0u;
        const unsigned delta = xor_reduce(term);
        if (delta == 0) {
          ++m;
        } else {
          const unsigned coef = F.exp[F.log[delta] + n - F.log[b]];  // delta / b
          unsigned bs = __shfl_up_sync(kAll, B, m & 31);
          if (lane < m || m >= 32) bs = 0;
          const unsigned nl = lam ^ F.mul(coef, bs);
          if (2 * L <= r) {
            B = lam;
            b = delta;
            L = r + 1 - L;
            m = 1;
          } else {
            ++m;
          }
          lam = nl;
        }
        if (L > p.t) break;
      }
      failed = L > p.t || L == 0;
      if (!failed) {
        lamv[lane] = static_cast<uint8_t>(lane <= L ? lam : 0);
        __syncwarp();
        // ---------------- roots: Lambda(alpha^-i) == 0  <=>  error at position i
        int found = 0;
        for (int i0 = 0; i0 < n; i0 += 32) {
          const int i = i0 + lane;
          unsigned v = 1;  // Lambda_0 = 1
          if (i < n) {
            const unsigned e = i ? n - i : 0;  // exponent of alpha^-i
            unsigned acc = 0;
            for (int k = 1; k <= L; ++k) {
              acc += e;
              if (acc >= static_cast<unsigned>(n)) acc -= n;
              v ^= F.mul_alpha(lamv[k], acc);
            }
          }
          const unsigned hit = __ballot_sync(kAll, i < n && v == 0);
          if (i < n && v == 0) {
            const int slot = found + __popc(hit & ((1u << lane) - 1u));
            if (slot < 32) pos[slot] = static_cast<uint8_t>(i);
          }
          found += __popc(hit);
        }
        __syncwarp();
        failed = found != L;
        if (!failed) {
          nerr = L;
          if (p.binary) {
            if (lane < L) word[pos[lane]] ^= 1;  // bch.h:80-83: every error value is 1
          } else {
            // ---------------- Forney: Omega = S * Lambda mod x^(2t);  e = X^(1-mu) Omega(1/X) / Lambda'(1/X)
            if (lane < L) {
              unsigned o = 0;
              for (int j = 0; j <= lane; ++j) o ^= F.mul(lamv[j], synd[lane - j]);
              omega[lane] = static_cast<uint8_t>(o);
            }
            __syncwarp();
            if (lane < L) {
              const unsigned pk = pos[lane];
              const unsigned e = pk ? n - pk : 0;  // 1/X = alpha^-pos
              unsigned num = 0, den = 0, acc = 0;  // acc = exponent of (1/X)^k
              for (int k = 0; k < L; ++k) {
                num ^= F.mul_alpha(omega[k], acc);
                if ((k & 1) == 0) den ^= F.mul_alpha(lamv[k + 1], acc);  // Lambda' = sum_{k odd} Lambda_k x^(k-1)
                acc += e;
                if (acc >= static_cast<unsigned>(n)) acc -= n;
              }
              // X^(1-mu): exponent pos * (1 - mu) mod n
              const unsigned xe = (pk * static_cast<unsigned>(((1 - p.mu) % n + n) % n)) % n;
              unsigned val = 0;
              if (den != 0 && num != 0) val = F.exp[(F.log[num] + n - F.log[den] + xe) % n];
              if (den == 0) val = 0;
              word[pk] ^= static_cast<uint8_t>(val);
              // a zero error value means the locator was not a true error locator
              failed = (val == 0);
            }
            failed = __any_sync(kAll, failed);
          }
          __syncwarp();
          // ---------------- cyclic.h:243-248: "Corrected word is not a codeword"
          if (!failed) failed = syndromes(p, F, word, s_rexp, synd, lane);
        }
      }
    }
    uint8_t *out = p.corrected + w * n;
    if (failed) {
      for (int i = lane; i < n; i += 32) out[i] = in[i];
    } else {
      for (int i = lane; i < n; i += 32) out[i] = word[i];
    }
    if (lane == 0) {
      p.failed[w] = failed ? 1 : 0;
      if (p.n_errors) p.n_errors[w] = static_cast<uint8_t>(failed ? 0 : nerr);
    }
    __syncwarp();
  }
}

}  // namespace

int gf_upload(const CodeSpec &spec, GfDevice *out) {
  *out = GfDevice();
  out->q = static_cast<int>(spec.q);
  out->n = static_cast<int>(spec.n);
  out->t = static_cast<int>(spec.t);
  out->nroots = static_cast<int>(spec.roots.size());
  out->binary = spec.family == 0;
  out->mu = spec.family == 0 ? 1 : static_cast<int>(spec.mu);
  out->step = spec.family == 0 ? 1 : static_cast<int>(spec.step);
  const size_t size = spec.F.size;
  std::vector<uint8_t> tab(3 * size);
  for (size_t i = 0; i < 2 * size; ++i) tab[i] = static_cast<uint8_t>(spec.F.exp[i]);
  for (size_t i = 0; i < size; ++i) tab[2 * size + i] = static_cast<uint8_t>(spec.F.log[i]);
  if (cudaMalloc(&out->tables, tab.size()) != cudaSuccess) return -1;
  if (cudaMemcpy(out->tables, tab.data(), tab.size(), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
  return 0;
}

void gf_free(GfDevice *d) {
  if (d->tables) cudaFree(d->tables);
  *d = GfDevice();
}

int gf_launch(const GfDevice &d, const uint8_t *words, uint64_t count, uint8_t *corrected, uint8_t *n_errors,
              uint8_t *failed, int sm_count, cudaStream_t stream) {
  if (!d.tables || d.t > 31 || d.nroots > 64 || d.step != 1) return -3;
  GfParams p{ d.q, d.n, d.t, d.nroots, d.binary, d.mu, d.tables, words, count, corrected, n_errors, failed };
  const size_t smem = 3 * (size_t(1) << d.q) + 64 + size_t(kGfWarps) * 512;
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gf_decode_kernel, kGfThreads, smem);
  const uint64_t want = (count + kGfWarps - 1) / kGfWarps;
  const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(want, uint64_t(std::max(1, occ)) * sm_count));
  gf_decode_kernel<<<grid, kGfThreads, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace ccgpu
