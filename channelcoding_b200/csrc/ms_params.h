// ms_params.h -- kernel parameter blocks shared by the min-sum kernels and the host API.
#pragma once
#include <cstdint>

namespace ccgpu {

constexpr int kMaxTaps = 128;      // row weight limit of the cyclic kernels
constexpr int kMsThreads = 128;    // threads per CTA of the cyclic kernels (4 warps)
constexpr int kCounterSlots = 8;   // ccgpu_counters as 8 x u64

enum : int { V_MS = 0, V_NMS = 1, V_OMS = 2, V_SCMS1 = 3, V_SCMS2 = 4, V_NMS2D = 5, V_SPA = 6,
             V_MS_Q = 7, V_NMS_Q = 8, V_OMS_Q = 9 /* fixed-point min-sum (ms_cyclic_q.cuh) */ };
enum : int { STOP_REF = 0, STOP_GF2 = 1, STOP_NONE = 2 };
enum : int { SRC_HBM = 0, SRC_PHILOX = 1, SRC_BITFLIP = 2 };

// Philox with the ten round keys precomputed on the host (they depend on the seed only): the rounds take their
// key from the constant bank as a LOP3 operand instead of two uniform-datapath adds per round
struct PhiloxKeys {
  uint32_t x[10], y[10];
};
inline PhiloxKeys philox_round_keys(uint64_t seed) {
  PhiloxKeys k;
  uint32_t kx = static_cast<uint32_t>(seed), ky = static_cast<uint32_t>(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    k.x[r] = kx;
    k.y[r] = ky;
    kx += 0x9E3779B9u;
    ky += 0xBB67AE85u;
  }
  return k;
}
// One launch of a min-sum kernel (passed as a __grid_constant__ kernel parameter).
struct MsParams {
  // ---- parity-check matrix: n columns, k rows (the cyclic kernels know the tap offsets at compile time)
  int32_t n, k;
  // ---- decoder
  int32_t variant, stop_rule, max_iter, src;
  int32_t stop_simple, pad1;  // 1: STOP_REF on an H where it is equivalent to "decided word is all-zero"
  float alpha_f, beta_f;  // alpha / beta converted double -> float exactly where the reference does
  double beta_d;          // OMS works in double (soft_decision.h:245-251)
  // ---- frame source
  const float *y;      // SRC_HBM: frames x n
  float sigma;         // SRC_PHILOX
  float llr_scale;     // SRC_PHILOX: y is multiplied by this (2 / sigma^2 for sum-product, else 1)
  uint32_t point;
  uint64_t seed, frame0;
  PhiloxKeys keys;     // SRC_PHILOX: the round keys of `seed` (channel.cuh)
  uint64_t frames;     // frames of this launch
  uint32_t flip_weight;  // SRC_BITFLIP: patterns of this weight, frame index = lexicographic rank
  // ---- outputs (nullable)
  uint8_t *bits;       // frames x n
  float *L;            // frames x n
  uint8_t *iter;       // frames
  uint8_t *failed;     // frames
  uint32_t *packed;    // compact layout: frames x ceil(n/32) words, bit (c & 31) of word (c >> 5) = decision of column c
  uint8_t *status;     // compact layout: frames, iteration index at which the stop test passed, 255 = failure
  unsigned long long *counters;  // kCounterSlots, accumulated with atomics
  unsigned long long *work;      // dynamic frame queue head, zeroed by the host before the launch
  unsigned work_batch;            // most frame indices a warp takes from the queue per atomic (>= 1)
  unsigned work_shift;            // guided schedule: a warp takes min(work_batch, remaining >> work_shift) frames
  int quick_hint;                 // host: enough all-positive frames are expected for the VN_QUICK kernels to pay
  // ---- fixed-point variants (V_*_Q): y_int = clamp(rint(y * q_scale), +-q_ymax); messages saturate at q_mmax;
  // fn_h(m) = max(rne(q_alpha * m / 1024) - q_beta, 0)
  float q_scale;
  int32_t q_ymax, q_mmax, q_alpha, q_beta;
  // the same constants as fp16x2 words (both halves equal), converted on the host so that a kernel that has to
  // rematerialise them inside its loop pays one constant-bank read, not an integer -> fp16 conversion
  uint32_t q_h2_mmax, q_h2_alpha /* q_alpha / 1024 */, q_h2_1024 /* 1024 */, q_h2_1024b /* 1024 + q_beta */;
};

// counter slots (ccgpu_counters layout)
enum : int { C_FRAMES = 0, C_FRAME_ERR = 1, C_BIT_ERR = 2, C_ITER = 3, C_FAIL = 4, C_UNDETECTED = 5 };

using ms_kernel_fn = void (*)(MsParams);

// vertical-node flavour a kernel is compiled for
enum : int { VN_PLAIN = 0 /* MS NMS OMS */, VN_SC = 1 /* SCMS1 SCMS2 */, VN_2D = 2 /* 2DNMS */, VN_SPA = 3 /* sum-product */,
             VN_QUICK = 4 /* + VN_PLAIN / VN_SC / VN_2D: the same with the all-positive-frame shortcut (ms_cyclic.cuh) */,
             VN_FIX = 7 /* fixed-point min-sum, two frames per lane (ms_cyclic_q.cuh) */, VN_COUNT = 8 };

struct MsCyclicEntry {
  const char *name;
  int n, k /* 0: rows at run time */, w, rpl, fpw, np, wrap, vn;
  int threads;  // CTA size
  int cta;      // 0: ms_cyclic_kernel (warp owns frames), 1: ms_cyclic_cta_kernel (CTA owns one frame),
                // 2: ms_cyclic_lane_kernel (lane owns a frame, small codes; kept in ccgpu_code::lane, not ::cyc)
  int slots;    // frames a lane / thread works on at once (2 for the fixed-point kernels, else 1)
  int dyn_smem; // cta == 1: bytes of dynamic shared memory the kernel needs (gather buffer of the float CTA kernel)
  const int *taps;
  ms_kernel_fn fn;
};

// registry filled by the instantiation units (ms_cyclic_inst.cu x CCGPU_GROUP)
int ms_cyclic_count();
const MsCyclicEntry *ms_cyclic_at(int i);

}  // namespace ccgpu
