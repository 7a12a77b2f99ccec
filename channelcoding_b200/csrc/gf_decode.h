// gf_decode.h -- host interface of the batched algebraic BCH/RS decoder (gf_decode.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "codes.hpp"

namespace ccgpu {

struct GfDevice {
  int q = 0, n = 0, t = 0, nroots = 0, binary = 0, mu = 1, step = 1;
  int recheck = 0;  // 1: always re-compute the syndromes of the corrected word (cyclic.h:243-248)
  uint8_t *tables = nullptr;  // exp[2*size] then log[size]
};

int gf_upload(const CodeSpec &spec, GfDevice *out);
void gf_free(GfDevice *d);
// words/corrected: count x n bytes; n_errors (nullable) / failed: count bytes; erasure_pos: count x
// max_erasures positions, erasure_cnt: count (both nullable).  0 ok, -1 CUDA error, -3 unsupported
int gf_launch(const GfDevice &d, const uint8_t *words, uint64_t count, const uint8_t *erasure_pos,
              const uint8_t *erasure_cnt, int max_erasures, uint8_t *corrected, uint8_t *n_errors, uint8_t *failed,
              int sm_count, cudaStream_t stream);

}  // namespace ccgpu
