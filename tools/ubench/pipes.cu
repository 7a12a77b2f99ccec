// pipes.cu -- which pipe do the candidate instructions of the packed (two frames per lane) min-sum kernel use,
// and at which rate?  Every op is timed alone and interleaved 1:1 with LOP3 (alu pipe) and FFMA (fma pipe):
// if the pair runs at the sum of the single rates the two are on different pipes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

enum Op { LOP3, FFMA, FADD, FADD2, FMNMX, FMNMX3, HADD2, HFMA2, HMNMX2, HMNMX2ABS, HSET2, VIADD2, VIMNMX2, VIMNMX32, VIADDMNMX2, PRMT, IMAD, FSEL, IADD3, HADD2ABS, LDS, STS, LDS64, LDS128, STS64, HSET2BF, LDSI, NOPS };
static const char *names[] = { "LOP3", "FFMA", "FADD", "FADD2", "FMNMX", "FMNMX3", "HADD2", "HFMA2", "HMNMX2", "HMNMX2|a|", "HSET2", "VIADD.16x2", "VIMNMX.S16x2", "VIMNMX3.S16x2", "VIADDMNMX.S16x2", "PRMT", "IMAD", "FSEL", "IADD3", "HADD2|a|", "LDS", "STS", "LDS.64", "LDS.128", "STS.64", "HSET2.BF|a|", "LDS(indep)" };

template <int OP> __device__ __forceinline__ void step(unsigned &a, unsigned b, unsigned c, unsigned long long &w, float *sm) {
  if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
  if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float *)&a) : "f"(__uint_as_float(b)), "f"(__uint_as_float(c)));
  if (OP == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(*(float *)&a) : "f"(__uint_as_float(b)));
  if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w) : "l"(((unsigned long long)b << 32) | c));
  if (OP == FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(*(float *)&a) : "f"(__uint_as_float(b)));
  if (OP == FMNMX3) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(*(float *)&a) : "f"(__uint_as_float(b)), "f"(__uint_as_float(c)));
  if (OP == HADD2) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
  if (OP == HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
  if (OP == HMNMX2) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
  if (OP == HMNMX2ABS) asm volatile("{.reg .b32 t; abs.f16x2 t, %1; min.f16x2 %0, %0, t;}" : "+r"(a) : "r"(b));
  if (OP == HADD2ABS) asm volatile("{.reg .b32 t; abs.f16x2 t, %1; add.rn.f16x2 %0, %0, t;}" : "+r"(a) : "r"(b));
  if (OP == HSET2) asm volatile("{.reg .b32 t; set.eq.u32.f16x2 t, %0, %1; xor.b32 %0, t, %2;}" : "+r"(a) : "r"(b), "r"(c));  // + one LOP3
  if (OP == VIADD2) asm volatile("add.s16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
  if (OP == VIMNMX2) asm volatile("min.s16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
  if (OP == VIMNMX32) a = __vimin3_s16x2(a, b, c);
  if (OP == VIADDMNMX2) a = __viaddmin_s16x2(a, b, c);
  if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(a) : "r"(b));
  if (OP == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
  if (OP == FSEL) asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; selp.b32 %0, %0, %1, p;}" : "+r"(a) : "r"(b), "r"(c));
  if (OP == IADD3) asm volatile("add.s32 %0, %0, %1;" : "+r"(a) : "r"(b));
  if (OP == LDS) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(sm) + (a & 0x7c))); a ^= __float_as_uint(v); }
  if (OP == LDS64) { unsigned v0, v1; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v0), "=r"(v1) : "r"((unsigned)__cvta_generic_to_shared(sm) + ((b & 0x3c) << 1))); a ^= v0 ^ v1; }
  if (OP == LDS128) { unsigned v0, v1, v2, v3; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"((unsigned)__cvta_generic_to_shared(sm) + ((b & 0x1c) << 2))); a ^= v0 ^ v1 ^ v2 ^ v3; }
  if (OP == LDSI) { unsigned v0; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v0) : "r"((unsigned)__cvta_generic_to_shared(sm) + (b & 0x7c))); a ^= v0; }
  if (OP == STS64) asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"((unsigned)__cvta_generic_to_shared(sm) + ((b & 0x3c) << 1)), "r"(a), "r"(c));
  if (OP == HSET2BF) asm volatile("{.reg .b32 t; abs.f16x2 t, %0; set.gt.f16x2.f16x2 %0, t, %1;}" : "+r"(a) : "r"(b));
  if (OP == STS) asm volatile("st.shared.f32 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(sm) + (b & 0x7c)), "f"(__uint_as_float(a)));
}

template <int A, int B> __global__ void __launch_bounds__(256) bench(unsigned *out, long long *cyc, unsigned b, unsigned c, int iters) {
  __shared__ __align__(16) float sm[64 * 8];
  unsigned x[8], z[8];
  unsigned long long w[8];
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 3 + i; z[i] = threadIdx.x + 7 * i; w[i] = x[i]; }
  sm[threadIdx.x] = 0; sm[threadIdx.x + 256] = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        step<A>(x[i], x[(i + 1) & 7], c, w[i], sm + (threadIdx.x & ~31));  // neighbour chain as operand: nothing folds
        if (B != NOPS) step<B>(z[i], z[(i + 1) & 7], c, w[i], sm + (threadIdx.x & ~31));
      }
  }
  const long long t1 = clock64();
  unsigned acc = 0;
  for (int i = 0; i < 8; ++i) acc ^= x[i] ^ z[i] ^ (unsigned)w[i] ^ (unsigned)(w[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int A, int B> double run(unsigned *out, long long *cyc, int ctas_per_sm) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 2000, grid = sms * ctas_per_sm;
  bench<A, B><<<grid, 256>>>(out, cyc, 0x3c003c00u, 0x00010001u, iters);
  bench<A, B><<<grid, 256>>>(out, cyc, 0x3c003c00u, 0x00010001u, iters);
  cudaDeviceSynchronize();
  long long h[4096]; cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
  const double instr = double(iters) * 32 * (B == NOPS ? 1 : 2) * 8 /*warps*/ * ctas_per_sm;  // warp instructions per SM
  return instr / mean;  // warp instructions per clock per SM
}
template <int A> void row(unsigned *out, long long *cyc) {
  const double alone = run<A, NOPS>(out, cyc, 4), wl = run<A, LOP3>(out, cyc, 4), wf = run<A, FFMA>(out, cyc, 4), wh = run<A, HADD2>(out, cyc, 4), ws = run<A, LDS>(out, cyc, 4);
  printf("%-16s alone %.2f   +LOP3 %.2f   +FFMA %.2f   +HADD2 %.2f   +LDS %.2f   (warp instr/clk/SM)\n", names[A], alone, wl, wf, wh, ws);
}
int main() {
  unsigned *out; long long *cyc;
  cudaMalloc(&out, 4096 * 256 * 4); cudaMalloc(&cyc, 4096 * 8);
  row<LOP3>(out, cyc); row<FFMA>(out, cyc); row<FADD>(out, cyc); row<FADD2>(out, cyc); row<FMNMX>(out, cyc); row<FMNMX3>(out, cyc);
  row<HADD2>(out, cyc); row<HADD2ABS>(out, cyc); row<HFMA2>(out, cyc); row<HMNMX2>(out, cyc); row<HMNMX2ABS>(out, cyc); row<HSET2>(out, cyc);
  row<VIADD2>(out, cyc); row<VIMNMX2>(out, cyc); row<VIMNMX32>(out, cyc); row<VIADDMNMX2>(out, cyc); row<PRMT>(out, cyc); row<IMAD>(out, cyc);
  row<FSEL>(out, cyc); row<IADD3>(out, cyc); row<LDS>(out, cyc); row<STS>(out, cyc);
  row<LDSI>(out, cyc); row<LDS64>(out, cyc); row<LDS128>(out, cyc); row<STS64>(out, cyc); row<HSET2BF>(out, cyc);
  return cudaGetLastError() != cudaSuccess;
}
