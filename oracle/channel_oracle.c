/* oracle/channel_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the device channel model (channelcoding_b200/csrc/channel.cuh):
 * all-zero codeword, BPSK 0 -> +1, y = 1 + sigma * z with
 * sigma = 1 / sqrt(2 * rate * 10^(EbN0/10))  -- reference src/simulation/simulation.c++:83-85,
 * :113-115, :125 (normal_distribution<float>(1.0, float(sigma)) on an all-zero word).
 *
 * The reference draws z from libstdc++'s mt19937_64 + normal_distribution, whose sequence is
 * implementation-defined; the engine uses counter-based Philox4x32-10 (Salmon et al., SC'11)
 * + Box-Muller instead, so noise parity with the reference is statistical only.
 * This file pins (a) the Philox integer stream bit-exactly (Random123 known answers in
 * tests/test_channel.py) and (b) the float transform within a stated tolerance (the device
 * uses MUFU-based __logf/__sincosf).
 */
#include <math.h>
#include <stdint.h>

#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void oracle_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int round = 0; round < 10; round++) {
    uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
    uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += PHILOX_W0;
    k1 += PHILOX_W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* counter = {frame_lo, frame_hi, block, point}, key = {seed_lo, seed_hi};
 * block b covers symbols 4b .. 4b+3:  (x0,x1) -> (z0,z1), (x2,x3) -> (z2,z3) with
 *   u = x0 * 2^-32 + 2^-33  in (0,1],  v = (int32)x1 * (pi * 2^-31) in [-pi, pi)
 *   rad = sqrt(-2 ln u);  z0 = rad * sin v,  z1 = rad * cos v */
static void box_muller(uint32_t x0, uint32_t x1, float *z0, float *z1) {
  float u = (float)x0 * 2.3283064365386963e-10f + 1.1641532182693481e-10f;
  float v = (float)(int32_t)x1 * 1.4629180792671596e-9f;
  float rad = sqrtf(-2.0f * logf(u));
  *z0 = rad * sinf(v);
  *z1 = rad * cosf(v);
}

double oracle_sigma(double rate, double ebno_db) {
  /* simulation.c++:83-85: 1.0f / sqrt(2 * rate * pow(10, ebno/10)) evaluated in double */
  return 1.0f / sqrt((2 * rate * pow(10, ebno_db / 10.0)));
}

void oracle_awgn_frame(uint64_t seed, uint32_t point, uint64_t frame, unsigned n, float sigma, float *y) {
  uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
  for (unsigned blk = 0; blk * 4 < n; blk++) {
    uint32_t ctr[4] = { (uint32_t)frame, (uint32_t)(frame >> 32), blk, point };
    uint32_t x[4];
    float z[4];
    oracle_philox4x32_10(ctr, key, x);
    box_muller(x[0], x[1], &z[0], &z[1]);
    box_muller(x[2], x[3], &z[2], &z[3]);
    for (unsigned j = 0; j < 4 && blk * 4 + j < n; j++) y[blk * 4 + j] = 1.0f + sigma * z[j];
  }
}

void oracle_awgn_batch(uint64_t seed, uint32_t point, uint64_t frame0, uint64_t frames, unsigned n,
                       float sigma, float *y) {
  for (uint64_t f = 0; f < frames; f++) oracle_awgn_frame(seed, point, frame0 + f, n, sigma, y + f * n);
}
