// Shared helpers for libb200match (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <stdlib.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libb200match is written for sm_100a (B200) only"
#endif

namespace b200m {

constexpr int kWarp = 32;

// Precision experiment (profiles/r02_single_product_experiment.md): the hi-x-hi-only code paths exist ONLY in a build made
// with EXTRA=-DB200M_SINGLE_EXPERIMENT.  As run-time branches in the MMA-issue and epilogue loops they had cost the
// product build 4 % of its throughput (the conv kernels 8.1 -> 8.9 ms per step), so the product build compiles them out.
#ifdef B200M_SINGLE_EXPERIMENT
constexpr bool kSingleExp = true;
#else
constexpr bool kSingleExp = false;
#endif

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t cdivz(size_t a, size_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return cdiv(a, b) * b; }

// ---- cp.async (LDGSTS) helpers --------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = pred ? 16 : 0;   // src-size 0 -> the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- warp reductions ------------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: one of these per kernel (function-local
// static) remembers the devices it was set on, so a second GPU in the same process gets its own opt-in.
struct SmemOptIn {
  unsigned long long done = 0;
  template <typename K>
  bool ensure(K kern, int bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (dev < 64 && ((done >> dev) & 1ull)) return true;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return false;
    if (dev < 64) done |= 1ull << dev;
    return true;
  }
};

// Launch bookkeeping: every launch goes through LAUNCH_CHECK so errors surface with a name and
// the handle's launch counter (bench.py "gpu_launches") stays truthful.
struct ProfRecord { const char* name; cudaEvent_t start, stop; };
struct Profiler {            // optional per-launch CUDA-event timing (bench.py roofline leg)
  bool enabled = false;
  ProfRecord* recs = nullptr;
  int n = 0, cap = 0;
};

struct LaunchCtx {
  cudaStream_t stream;
  long long* counter;
  const char** err_where;
  cudaError_t err;
  Profiler* prof;
  bool pdl = false;          // launch_pdl attaches the programmatic-serialisation attribute (small batches only, see api.cu)
};

// RAII: brackets one kernel launch with events on the launching stream when profiling is on.
struct ProfScope {
  Profiler* p; cudaStream_t s; int idx;
  ProfScope(LaunchCtx& ctx, const char* name) : p(ctx.prof), s(ctx.stream), idx(-1) {
    if (p && p->enabled && p->n < p->cap) {
      idx = p->n++;
      p->recs[idx].name = name;
      cudaEventCreate(&p->recs[idx].start);
      cudaEventCreate(&p->recs[idx].stop);
      cudaEventRecord(p->recs[idx].start, s);
    }
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(p->recs[idx].stop, s); }
};

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------
// Every kernel of the hot path starts with pdl_trigger() (the NEXT kernel of the stream may be scheduled as soon as all
// CTAs of this one have started: its launch latency, CTA rasterisation and barrier / tensor-memory set-up overlap this
// kernel's tail) and executes pdl_wait() before it touches global memory a predecessor may have written -- or writes
// anything a predecessor may still read.  pdl_wait() returns once the preceding kernel has COMPLETED and its writes are
// visible; as every kernel of the chain waits, completion is transitive (kernel k+1 cannot finish before kernel k).
// Both are no-ops in a kernel launched without the attribute.  The attribute is attached for SMALL batches only
// (LaunchCtx::pdl, set per call in api.cu): measured on B200 inside the replayed CUDA graph, one pair per call gains 3-4 %
// (1.94 -> 1.87 ms), while 64 pairs per call LOSE 0.9 % (26.83 -> 27.08 ms: the graph's kernel-to-kernel gaps are already
// ~1 us and the kernels run for 0.1-3 ms).  B200M_PDL=0 in the environment launches everything fully serialised.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  static const int on = [] {
    const char* e = getenv("B200M_PDL");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return on != 0;
}

// Kernel classes for the large-batch PDL mask: which classes keep the attribute when LaunchCtx::pdl is off.  Default: the
// Sinkhorn iterations only -- the next iteration's CTAs request their first rows of S while the last CTA of each pair
// still folds the partial column sums (64 pairs per call: 27.17 -> 27.05 ms per step, twice on the same box; conv,
// attention + fused layer, post-processing classes each measured neutral or slower).  B200M_PDL_MASK overrides.
enum PdlClass { kPdlConv = 1, kPdlAttn = 2, kPdlGnn = 4, kPdlGemm = 8, kPdlOt = 16, kPdlPost = 32 };
inline int pdl_large_mask() {
  static const int m = [] {
    const char* e = getenv("B200M_PDL_MASK");
    return e ? atoi(e) : (int)kPdlOt;
  }();
  return m;
}

// kern<<<grid, block, smem, stream>>>(args...) with the programmatic-stream-serialisation attribute
template <typename... KArgs, typename... Args>
inline void launch_pdl(LaunchCtx& ctx, int cls, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx.stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (pdl_enabled() && (ctx.pdl || (pdl_large_mask() & cls))) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#define B200M_LAUNCH_CHECK(ctx, name)                         \
  do {                                                        \
    cudaError_t e__ = cudaGetLastError();                     \
    if ((ctx).counter) ++*(ctx).counter;                      \
    if (e__ != cudaSuccess && (ctx).err == cudaSuccess) {     \
      (ctx).err = e__;                                        \
      if ((ctx).err_where) *(ctx).err_where = name;           \
    }                                                         \
  } while (0)

}  // namespace b200m
