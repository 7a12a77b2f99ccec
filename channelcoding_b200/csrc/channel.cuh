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

// two uniforms -> two N(0,1):  u in (0,1], v in [-pi,pi)  (MUFU lg2 / sin / cos)
__device__ __forceinline__ float2 box_muller(uint32_t x0, uint32_t x1) {
  const float u = __fmaf_rn(static_cast<float>(x0), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  const float v = static_cast<float>(static_cast<int32_t>(x1)) * 1.4629180792671596e-9f;
  const float rad = sqrtf(-2.0f * __logf(u));
  float s, c;
  __sincosf(v, &s, &c);
  return make_float2(rad * s, rad * c);
}

// the four channel values y = 1 + sigma * z of block `blk` of frame `frame`
__device__ __forceinline__ float4 awgn_block(uint64_t seed, uint32_t point, uint64_t frame, uint32_t blk,
                                             float sigma) {
  const uint4 x = philox4x32_10(make_uint4(static_cast<uint32_t>(frame), static_cast<uint32_t>(frame >> 32), blk, point),
                                make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  const float2 a = box_muller(x.x, x.y);
  const float2 b = box_muller(x.z, x.w);
  return make_float4(__fadd_rn(1.0f, __fmul_rn(sigma, a.x)), __fadd_rn(1.0f, __fmul_rn(sigma, a.y)),
                     __fadd_rn(1.0f, __fmul_rn(sigma, b.x)), __fadd_rn(1.0f, __fmul_rn(sigma, b.y)));
}

}  // namespace ccgpu
