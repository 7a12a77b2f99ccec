// ms_csr.h -- host interface of the general parity-check-matrix min-sum kernel (ms_csr.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "ms_params.h"

namespace ccgpu {

// device-resident description of an arbitrary 0/1 matrix H (rows x n)
struct MsCsrDevice {
  int rows = 0, n = 0, wmax = 0, edges = 0;
  uint16_t *ell_col = nullptr;   // [wmax][rows]  column of slot j of row r, 0xffff = empty
  uint16_t *csc_ptr = nullptr;   // [n + 1]
  uint16_t *csc_edge = nullptr;  // [edges]       ELL slot index (j * rows + r), rows ascending per column
  size_t smem_bytes = 0;
  int threads = 0;
};

// 0 ok, -1 CUDA error, -3 does not fit the kernel's shared-memory layout
int ms_csr_upload(const uint8_t *H, unsigned rows, unsigned n, MsCsrDevice *out);
void ms_csr_free(MsCsrDevice *d);
int ms_csr_launch(const MsCsrDevice &d, const MsParams &mp, int sm_count, cudaStream_t stream);

}  // namespace ccgpu
