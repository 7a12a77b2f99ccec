// ms_params.h -- kernel parameter blocks shared by the min-sum kernels and the host API.
#pragma once
#include <cstdint>

namespace ccgpu {

constexpr int kMaxTaps = 128;      // row weight limit of the cyclic kernels
constexpr int kMsThreads = 128;    // threads per CTA of the cyclic kernels (4 warps)
constexpr int kCounterSlots = 8;   // ccgpu_counters as 8 x u64

enum : int { V_MS = 0, V_NMS = 1, V_OMS = 2, V_SCMS1 = 3, V_SCMS2 = 4, V_NMS2D = 5, V_SPA = 6 };
enum : int { STOP_REF = 0, STOP_GF2 = 1, STOP_NONE = 2 };
enum : int { SRC_HBM = 0, SRC_PHILOX = 1, SRC_BITFLIP = 2 };

// One launch of a min-sum kernel.  Passed as a __grid_constant__ kernel parameter, so the tap
// offsets below live in the constant bank and fold into instruction operands after unrolling.
struct MsParams {
  // ---- parity-check structure: row r has ones at (r + tap[j]) (mod n if wrap), j < w
  int32_t n, k, w, fpw;  // columns, rows, row weight, frames per warp
  int16_t tap[kMaxTaps];
  // ---- decoder
  int32_t variant, stop_rule, max_iter, src;
  float alpha_f, beta_f;  // alpha / beta converted double -> float exactly where the reference does
  double beta_d;          // OMS works in double (soft_decision.h:245-251)
  // ---- frame source
  const float *y;      // SRC_HBM: frames x n
  float sigma;         // SRC_PHILOX
  uint32_t point;
  uint64_t seed, frame0;
  uint64_t frames;     // frames of this launch
  uint32_t flip_weight;  // SRC_BITFLIP: patterns of this weight, frame index = lexicographic rank
  uint32_t pad0;
  // ---- outputs (nullable)
  uint8_t *bits;       // frames x n
  float *L;            // frames x n
  uint8_t *iter;       // frames
  uint8_t *failed;     // frames
  unsigned long long *counters;  // kCounterSlots, accumulated with atomics
};

// counter slots (ccgpu_counters layout)
enum : int { C_FRAMES = 0, C_FRAME_ERR = 1, C_BIT_ERR = 2, C_ITER = 3, C_FAIL = 4, C_UNDETECTED = 5 };

using ms_kernel_fn = void (*)(MsParams);

struct MsCyclicEntry {
  int w, rpl, np, sc, wrap;
  ms_kernel_fn fn;
  const char *name;
};

// registry filled by the instantiation units (ms_cyclic_inst_*.cu)
const MsCyclicEntry *ms_cyclic_find(int w, int rpl, int np, int sc, int wrap);
int ms_cyclic_count();
const MsCyclicEntry *ms_cyclic_at(int i);

}  // namespace ccgpu
