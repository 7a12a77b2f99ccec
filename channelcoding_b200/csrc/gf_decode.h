// gf_decode.h -- host interface of the batched algebraic BCH/RS decoder (gf_decode.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "codes.hpp"

namespace ccgpu {

struct GfDevice {
  int q = 0, n = 0, t = 0, nroots = 0, binary = 0, mu = 1, step = 1;
  uint8_t *tables = nullptr;  // exp[2*size] then log[size]
};

int gf_upload(const CodeSpec &spec, GfDevice *out);
void gf_free(GfDevice *d);
// words/corrected: count x n bytes; n_errors (nullable) / failed: count bytes.  0 ok, -1 error
int gf_launch(const GfDevice &d, const uint8_t *words, uint64_t count, uint8_t *corrected, uint8_t *n_errors,
              uint8_t *failed, int sm_count, cudaStream_t stream);

}  // namespace ccgpu
