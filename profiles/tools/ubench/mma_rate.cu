// microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16) as a function of N and of the A source (smem / tmem)
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../image_matching_b200/csrc/tc_common.cuh"
using namespace b200m::tc;
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int N, bool TS, int NACC>
__global__ void __launch_bounds__(128, 2) k(long long* out, int iters, int cols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (threadIdx.x < 32) tmem_alloc(&slot, cols);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = instr_desc(0, 128, N);
    const uint64_t a = smem_desc_sw128(smem_u32(smem)), b = smem_desc_sw128(smem_u32(smem) + 32768);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t d = tb + 64 + (u % NACC) * N;     // NACC independent accumulators (dependent chain when NACC == 1)
        if (TS) mma_ts(d, tb + (u & 3) * 8, b + 2 * (u & 3), idesc, 1);
        else mma_bf16(d, a + 2 * (u & 3), b + 2 * (u & 3), idesc, 1);
      }
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tb, cols); }
}
template <int N, bool TS, int NW>
__global__ void __launch_bounds__(128, 2) kw(long long* out, int iters, int cols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  fence_proxy_async();
  if (threadIdx.x < 32) tmem_alloc(&slot, cols);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < NW) {
    const uint32_t idesc = instr_desc(0, 128, N);
    const uint64_t a = smem_desc_sw128(smem_u32(smem)), b = smem_desc_sw128(smem_u32(smem) + 32768);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t d = tb + 64 + w * N;
        if (TS) mma_ts(d, tb + (u & 3) * 8, b + 2 * (u & 3), idesc, 1);
        else mma_bf16(d, a + 2 * (u & 3), b + 2 * (u & 3), idesc, 1);
      }
    }
    tc_commit(&bar[w]);
    mbar_wait(&bar[w], 0);
    long long t1 = clock64();
    if (w == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tb, cols); }
}
template <int N, bool TS, int NW>
void runw(const char* name, int ctas = 148) {
  long long* d; cudaMalloc(&d, 296 * 8);
  auto kern = kw<N, TS, NW>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  const int iters = 2000;
  kern<<<ctas, 128, 65536 + 1024>>>(d, iters, ctas > 148 ? 256 : 512);
  kern<<<ctas, 128, 65536 + 1024>>>(d, iters, ctas > 148 ? 256 : 512);
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("%s ctas=%d issuing warps=%d N=%3d: %.1f clk per MMA per warp -> %.1f per CTA (ideal %.1f)  %s\n", name, ctas, NW, N, (double)h[0] / (iters * 8.0), (double)h[0] / (iters * 8.0) / NW, 128.0 * N * 16 * 2 / 8192, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}
template <int N, bool TS, int NACC = 1>
void run(const char* name, int ctas = 148) {
  long long* d; cudaMalloc(&d, 296 * 8);
  auto kern = k<N, TS, NACC>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  const int iters = 2000;
  kern<<<ctas, 128, 65536 + 1024>>>(d, iters, ctas > 148 ? 256 : 512);
  kern<<<ctas, 128, 65536 + 1024>>>(d, iters, ctas > 148 ? 256 : 512);
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("%s ctas=%d NACC=%d N=%3d: %.1f clk per MMA (ideal %.1f)  %s\n", name, ctas, NACC, N, (double)h[0] / (iters * 8.0), 128.0 * N * 16 * 2 / 8192, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  runw<32, true, 1>("TS"); runw<32, true, 2>("TS"); runw<32, true, 4>("TS"); runw<32, true, 2>("TS", 296);
  runw<48, true, 2>("TS"); runw<48, true, 2>("TS", 296);
  runw<64, false, 1>("SS"); runw<64, false, 2>("SS"); runw<64, false, 2>("SS", 296);
  runw<64, true, 2>("TS"); runw<64, true, 2>("TS", 296);
  return 0;
}
