// instruction throughput per SM (lanes per clock) for the softmax inner-loop instructions on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
template <int OP>
__global__ void __launch_bounds__(1024, 1) k(float* out, long long* cyc, int iters, float seed) {
  float a[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 1e-3f + i; h[i] = __float_as_uint(a[i]); }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(a[(i + 1) & 7])); a[i] = __uint_as_float(h[i] | 0x3f000000u); }
      if (OP == 9) { a[i] = __uint_as_float(__float_as_uint(a[i]) | 0x3f000000u); }
      if (OP == 2) { unsigned short hh = (unsigned short)h[i]; const unsigned short n1 = 0xBC00;
                     asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(a[i]) : "h"(hh), "h"(n1)); }
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(1.0001f), "f"(0.5f));
      if (OP == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(a[(i + 2) & 7]));
      if (OP == 5) asm volatile("{.reg .b32 t; max.f16x2 t, %0, %1; max.f16x2 %0, t, %2;}" : "+r"(h[i]) : "r"(h[(i + 1) & 7]), "r"(h[(i + 2) & 7]));
      if (OP == 6) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(1.5f));
      if (OP == 7) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(h[i]) : "r"(h[(i + 1) & 7]), "r"(0x12345u));
      if (OP == 10) { unsigned long long v = ((unsigned long long)__float_as_uint(a[(i + 1) & 7]) << 32) | __float_as_uint(a[i]);
                      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(0x3f8003473f800347ull), "l"(0x3f0000003f000000ull));
                      a[i] = __uint_as_float((uint32_t)v); if (i == 7) a[0] += __uint_as_float((uint32_t)(v >> 32)) * 1e-30f; }
      if (OP == 8) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(a[(i + 1) & 7])); a[i] = __uint_as_float(h[i] | 0x3f000000u); }
    }
  }
  long long t1 = clock64();
  float s = 0; uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += a[i]; x ^= h[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(x);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name) {
  float* o; long long* c; cudaMalloc(&o, 148 * 1024 * 4); cudaMalloc(&c, 148 * 8);
  const int iters = 4000;
  k<OP><<<148, 1024>>>(o, c, iters, 0.1f);
  k<OP><<<148, 1024>>>(o, c, iters, 0.1f);
  long long h[148]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-28s %.1f lanes/clk/SM\n", name, 1024.0 * 8 * iters / (double)h[0]);
  cudaFree(o); cudaFree(c);
}
int main() {
  run<0>("MUFU.EX2"); run<9>("LOP (baseline for F2FP test)"); run<1>("F2FP.F16.F32.PACK_AB (per instr)"); run<8>("F2FP.BF16 (per instr)"); run<2>("FHFMA"); run<3>("FFMA"); run<4>("FMNMX3");
  run<10>("FFMA2 (fma.rn.f32x2, per instr)"); run<5>("HMNMX2 x2 / VHMNMX"); run<6>("FADD"); run<7>("LOP3");
  return 0;
}
