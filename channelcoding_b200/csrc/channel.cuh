// channel.cuh -- BPSK over AWGN, counter-based (device side).
//
// Replaces the noise source of the reference's Monte-Carlo loop
//   std::normal_distribution<float>(1.0, float(sigma)) on std::mt19937_64
//   (reference src/simulation/simulation.c++:113-115, :125; all-zero codeword, BPSK 0 -> +1)
// with Philox4x32-10 (Salmon et al., SC'11) + Box-Muller so that any frame of any Eb/N0 point can
// be generated independently on any GPU:  counter = {frame_lo, frame_hi, block, point},
// key = {seed_lo, seed_hi};  block b yields the four symbols 4b..4b+3 of that frame.
// oracle/channel_oracle.c restates this on the CPU (Random123 known answers pin the integer part).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "ms_params.h"

namespace ccgpu {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKeys &k) {
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x[round], lo1, hi0 ^ c.w ^ k.y[round], lo0);
  }
  return c;
}

// two uniforms -> two N(0,1):  u in (0,1], v in [-pi,pi).  Everything on the MUFU unit: lg2 / sqrt / sin / cos
// (u >= 2^-33 is never subnormal, so the .ftz forms are exact equivalents and save the range fix-up; sqrt.approx(0) = 0
// covers u == 1).  Relative error of a sample about 1e-6, far below the Monte-Carlo resolution; the CPU restatement
// (oracle/channel_oracle.c, libm) is compared with a tolerance (tests/test_gpu_parity.py::test_channel_kernel).
__device__ __forceinline__ float2 box_muller(uint32_t x0, uint32_t x1) {
  const float u = __fmaf_rn(static_cast<float>(x0), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  const float v = static_cast<float>(static_cast<int32_t>(x1)) * 1.4629180792671596e-9f;
  float lg, rad;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(lg * -1.3862943611198906f));  // -2 ln 2 * lg2 u = -2 ln u
  float s, c;
  __sincosf(v, &s, &c);
  return make_float2(rad * s, rad * c);
}

// the four channel values y = 1 + sigma * z of block `blk` of frame `frame`
__device__ __forceinline__ float4 awgn_from_bits(const uint4 x, float sigma) {
  const float2 a = box_muller(x.x, x.y);
  const float2 b = box_muller(x.z, x.w);
  return make_float4(__fmaf_rn(sigma, a.x, 1.0f), __fmaf_rn(sigma, a.y, 1.0f), __fmaf_rn(sigma, b.x, 1.0f),
                     __fmaf_rn(sigma, b.y, 1.0f));
}
__device__ __forceinline__ float4 awgn_block(uint64_t seed, uint32_t point, uint64_t frame, uint32_t blk,
                                             float sigma) {
  return awgn_from_bits(philox4x32_10(make_uint4(static_cast<uint32_t>(frame), static_cast<uint32_t>(frame >> 32), blk, point),
                                      make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32))),
                        sigma);
}
__device__ __forceinline__ float4 awgn_block(const PhiloxKeys &keys, uint32_t point, uint64_t frame, uint32_t blk,
                                             float sigma) {
  return awgn_from_bits(philox4x32_10(make_uint4(static_cast<uint32_t>(frame), static_cast<uint32_t>(frame >> 32), blk, point), keys),
                        sigma);
}

}  // namespace ccgpu
