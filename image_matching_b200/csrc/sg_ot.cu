// Log-domain Sinkhorn optimal transport and mutual-nearest-neighbour match selection.
// Reference: superglue/models/superglue_test.py:141-170 (log_sinkhorn_iterations, log_optimal_transport),
// :268-285 (match selection).  HBM/L2-bound streaming kernels: the (N+1)x(M+1) couplings matrix is never
// materialised -- the bin row/column are the scalar alpha -- and every sweep re-reads S (L2-resident when
// the caller micro-batches pairs).  LSE uses the same max-shifted two-pass form as torch.logsumexp.
#include "kernels.cuh"

namespace b200m {

struct PairDims { int n, m; float norm, mu_bin, nu_bin; };

__device__ __forceinline__ PairDims pair_dims(const OtParams& p, int b) {
  PairDims d;
  d.n = p.counts0 ? p.counts0[b] : p.N;
  d.m = p.counts1 ? p.counts1[b] : p.M;
  d.norm = -logf((float)d.m + (float)d.n);          // norm = -(ms + ns).log()
  d.mu_bin = logf((float)d.m) + d.norm;             // log_mu[-1] = ns.log() + norm  (ns = #columns)
  d.nu_bin = logf((float)d.n) + d.norm;             // log_nu[-1] = ms.log() + norm  (ms = #rows)
  return d;
}

__global__ void ot_init_kernel(float* u, float* v, int total) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) { u[i] = 0.f; v[i] = 0.f; }
}

void launch_ot_init(LaunchCtx& ctx, const OtParams& p) {
  ProfScope prof__(ctx, "ot_init");
  int total = p.B * p.ld_uv;
  ot_init_kernel<<<cdiv(total, 256), 256, 0, ctx.stream>>>(p.u, p.v, total);
  B200M_LAUNCH_CHECK(ctx, "ot_init");
}

// u = log_mu - logsumexp(Z + v, dim=2): one warp per row (row n is the dustbin row).
__global__ void __launch_bounds__(256) ot_row_kernel(OtParams p) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const PairDims d = pair_dims(p, b);
  if (i > d.n || d.n == 0 || d.m == 0) return;
  const float* v = p.v + (size_t)b * p.ld_uv;
  const float* srow = p.S + (size_t)b * p.strideS + (size_t)i * p.ldS;
  const bool bin_row = (i == d.n);
  float mx = -INFINITY;
  for (int j = lane; j <= d.m; j += 32) {
    float c = (bin_row || j == d.m) ? p.alpha : srow[j];
    mx = fmaxf(mx, c + v[j]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j <= d.m; j += 32) {
    float c = (bin_row || j == d.m) ? p.alpha : srow[j];
    sum += expf((c + v[j]) - mx);
  }
  sum = warp_sum(sum);
  if (lane == 0) p.u[(size_t)b * p.ld_uv + i] = (bin_row ? d.mu_bin : d.norm) - (logf(sum) + mx);
}

void launch_ot_row_update(LaunchCtx& ctx, const OtParams& p) {
  ProfScope prof__(ctx, "ot_row_update");
  dim3 grid(cdiv(p.N + 1, 8), p.B);
  ot_row_kernel<<<grid, 256, 0, ctx.stream>>>(p);
  B200M_LAUNCH_CHECK(ctx, "ot_row");
}

// v = log_nu - logsumexp(Z + u, dim=1): a block owns 32 columns, its 8 warps stride over the rows with
// coalesced 128 B row reads; per-column partial max / sum are combined through shared memory.
__global__ void __launch_bounds__(256) ot_col_kernel(OtParams p) {
  __shared__ float red[8][33];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  const PairDims d = pair_dims(p, b);
  if (d.n == 0 || d.m == 0) return;               // uniform per block
  if (blockIdx.x * 32 > d.m) return;              // uniform per block
  const float* u = p.u + (size_t)b * p.ld_uv;
  const float* S = p.S + (size_t)b * p.strideS;
  const bool active = j <= d.m;
  const bool bin_col = (j == d.m);
  // one streaming pass with a running (max, sum) pair per column -- halves the L2 traffic of the column sweep;
  // the partial pairs of the 8 warps are merged exactly like a flash-softmax split
  float mx = -INFINITY, sum = 0.f;
  if (active)
    for (int i = w; i <= d.n; i += 8) {
      const float c = (bin_col || i == d.n) ? p.alpha : S[(size_t)i * p.ldS + j];
      const float x = c + u[i];
      if (x > mx) { sum = sum * expf(mx - x) + 1.f; mx = x; }
      else sum += expf(x - mx);
    }
  __shared__ float reds[8][33];
  red[w][lane] = mx;
  reds[w][lane] = sum;
  __syncthreads();
  if (w == 0 && active) {
    float M = red[0][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) M = fmaxf(M, red[k][lane]);
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += red[k][lane] == -INFINITY ? 0.f : reds[k][lane] * expf(red[k][lane] - M);
    p.v[(size_t)b * p.ld_uv + j] = (bin_col ? d.nu_bin : d.norm) - (logf(tot) + M);
  }
}

void launch_ot_col_update(LaunchCtx& ctx, const OtParams& p) {
  ProfScope prof__(ctx, "ot_col_update");
  dim3 grid(cdiv(p.M + 1, 32), p.B);
  ot_col_kernel<<<grid, 256, 0, ctx.stream>>>(p);
  B200M_LAUNCH_CHECK(ctx, "ot_col");
}

// ---------------------------------------------------------------------------------------------------------------
// Fused Sinkhorn iteration for M <= 1024 columns: ONE launch per iteration, S is read once per iteration (the
// unfused pair of kernels above reads it three times), two exps per element.
//   A CTA (8 warps) owns 64 consecutive rows of one pair.
//   prologue : v of the previous iteration is rebuilt from the per-CTA column partials the previous launch wrote
//              (<= 17 partial rows per pair, flash-softmax style merge), so no separate combine launch is needed.
//   rows     : a warp owns 8 rows.  A row (<= 1024 floats) lives in 32 registers per lane (8 coalesced float4 loads):
//              row LSE with v (two-pass in registers, one exp per element) -> u_i; then y = c + u_i is folded into the
//              warp's running per-column (max, sum) state, also in registers (branch-free online LSE: one exp per
//              element, exp(-|y - m|), because one of the two rescale factors is always 1).
//   epilogue : the 8 warps' column states are tree-merged through shared memory and written as ONE partial row.
// ot_finish_kernel turns the last partials into v for the match selection.
constexpr int kOtRowsPerWarp = 8;
constexpr int kOtRowsPerCta = 8 * kOtRowsPerWarp;
constexpr int kOtFusedMaxM = 1024;

int ot_fused_parts(int N) { return cdiv(N + 1, kOtRowsPerCta); }

constexpr float kOtNegBig = -1.0e30f;    // "empty" sentinel: finite, so no inf - inf can appear in the branch-free updates
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ot_ex2(float x) {    // 2^x, one MUFU op (ex2.approx.ftz: rel. error 2^-22)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// (m, s) <- (m, s) (+) y : running logsumexp state, branch-free, one exp (one of the two rescale factors is 1)
__device__ __forceinline__ void lse_push(float& m, float& s, float y) {
  const float dlt = y - m;
  const float e = ot_ex2(-fabsf(dlt) * kLog2e);
  s = dlt > 0.f ? fmaf(s, e, 1.f) : s + e;
  m = fmaxf(m, y);
}
// (m, s) <- (m, s) (+) (m2, s2); an empty state is (kOtNegBig, 0)
__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
  const float M = fmaxf(m, m2);
  s = s * ot_ex2((m - M) * kLog2e) + s2 * ot_ex2((m2 - M) * kLog2e);
  m = M;
}

__global__ void __launch_bounds__(256, 2) ot_iter_kernel(OtParams p, const float2* __restrict__ part_in,
                                                        float2* __restrict__ part_out, int max_parts, int ld_part,
                                                        int first) {
  __shared__ __align__(16) float v_s[kOtFusedMaxM + 4];
  __shared__ __align__(16) float2 xch[4][kOtFusedMaxM + 4];      // warp-state exchange of the tree merge (32 KB)
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const PairDims d = pair_dims(p, b);
  if (d.n == 0 || d.m == 0) return;                       // uniform per block
  const int parts = cdiv(d.n + 1, kOtRowsPerCta);
  if ((int)blockIdx.x >= parts) return;                   // uniform per block
  // ---- prologue: v_j = log_nu_j - logsumexp_i(c_ij + u_i) from the previous launch's partials (v = 0 at the start)
  // (thread = two adjacent columns; all partial rows of a batch are fetched before any is merged, so the L2 round trips
  // overlap instead of forming a dependent chain)
  constexpr int kBatch = 18;
  for (int j0 = 2 * threadIdx.x; j0 < kOtFusedMaxM + 4; j0 += 512) {
    float va = 0.f, vb = 0.f;
    if (!first && j0 <= d.m) {
      const float4* col = reinterpret_cast<const float4*>(part_in + (size_t)b * max_parts * ld_part + j0);
      const size_t ldq = (size_t)(ld_part >> 1);
      float m0 = kOtNegBig, s0 = 0.f, m1 = kOtNegBig, s1 = 0.f;
      for (int q0 = 0; q0 < parts; q0 += kBatch) {
        float4 t[kBatch];
#pragma unroll
        for (int q = 0; q < kBatch; ++q)
          t[q] = q0 + q < parts ? __ldg(col + (size_t)(q0 + q) * ldq) : make_float4(kOtNegBig, 0.f, kOtNegBig, 0.f);
#pragma unroll
        for (int q = 0; q < kBatch; ++q) {
          lse_merge(m0, s0, t[q].x, t[q].y);
          lse_merge(m1, s1, t[q].z, t[q].w);
        }
      }
      va = (j0 == d.m ? d.nu_bin : d.norm) - (logf(s0) + m0);
      if (j0 + 1 <= d.m) vb = (j0 + 1 == d.m ? d.nu_bin : d.norm) - (logf(s1) + m1);
    }
    v_s[j0] = va;
    v_s[j0 + 1] = vb;
  }
  __syncthreads();
  const int row0 = blockIdx.x * kOtRowsPerCta + warp * kOtRowsPerWarp;
  const float xbin = p.alpha + v_s[d.m];                  // dustbin column term of every row LSE
  // columns of this lane: 4*lane + 128*k + {0..3}.  Groups k < kfull are valid in every lane, group kfull is the ragged
  // one (columns >= m masked with the sentinel), groups above it are skipped (warp-uniform conditions).
  const int kfull = d.m >> 7;
  float cm[32], cs[32];
#pragma unroll
  for (int q = 0; q < 32; ++q) { cm[q] = kOtNegBig; cs[q] = 0.f; }
  float bm = kOtNegBig, bs = 0.f;                         // dustbin column state (same in every lane)
  const float* S = p.S + (size_t)b * p.strideS;
  float* u = p.u + (size_t)b * p.ld_uv;
  const int row_end = min(row0 + kOtRowsPerWarp, d.n + 1);
  for (int i = row0; i < row_end; ++i) {
    const bool bin_row = (i == d.n);
    const float* srow = S + (size_t)i * p.ldS + 4 * lane;
    float c[32];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k <= kfull) {
        const int j = 4 * lane + 128 * k;
        float4 t = make_float4(p.alpha, p.alpha, p.alpha, p.alpha);
        if (!bin_row && j < d.m) t = __ldg(reinterpret_cast<const float4*>(srow + 128 * k));   // rows padded to ldS
        if (k == kfull) {
          t.x = j < d.m ? t.x : kOtNegBig;
          t.y = j + 1 < d.m ? t.y : kOtNegBig;
          t.z = j + 2 < d.m ? t.z : kOtNegBig;
          t.w = j + 3 < d.m ? t.w : kOtNegBig;
        }
        c[4 * k] = t.x; c[4 * k + 1] = t.y; c[4 * k + 2] = t.z; c[4 * k + 3] = t.w;
      }
    }
    if (i + 1 < row_end && i + 1 < d.n) {   // pull the next row towards L1 while this one is processed
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k <= kfull && 4 * lane + 128 * k < d.m)
          asm volatile("prefetch.global.L1 [%0];" ::"l"(srow + p.ldS + 128 * k));
    }
    // ---- u_i = log_mu_i - logsumexp_j(c_ij + v_j), dustbin column included
    float mx = xbin;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k <= kfull) {
        const float4 vv = *reinterpret_cast<const float4*>(v_s + 4 * lane + 128 * k);
        mx = fmaxf(mx, fmaxf(fmaxf(c[4 * k] + vv.x, c[4 * k + 1] + vv.y), fmaxf(c[4 * k + 2] + vv.z, c[4 * k + 3] + vv.w)));
      }
    mx = warp_max(mx);
    const float nmx = -mx * kLog2e;
    float sum = lane == 0 ? ot_ex2(fmaf(xbin, kLog2e, nmx)) : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k <= kfull) {
        const float4 vv = *reinterpret_cast<const float4*>(v_s + 4 * lane + 128 * k);
        sum += ot_ex2(fmaf(c[4 * k] + vv.x, kLog2e, nmx)) + ot_ex2(fmaf(c[4 * k + 1] + vv.y, kLog2e, nmx));
        sum += ot_ex2(fmaf(c[4 * k + 2] + vv.z, kLog2e, nmx)) + ot_ex2(fmaf(c[4 * k + 3] + vv.w, kLog2e, nmx));
      }
    sum = warp_sum(sum);
    const float ui = (bin_row ? d.mu_bin : d.norm) - (logf(sum) + mx);
    if (lane == 0) u[i] = ui;
    // ---- fold this row into the column states: y_ij = c_ij + u_i (masked columns accumulate garbage that is never read)
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k <= kfull) {
#pragma unroll
        for (int e = 0; e < 4; ++e) lse_push(cm[4 * k + e], cs[4 * k + e], c[4 * k + e] + ui);
      }
    lse_push(bm, bs, p.alpha + ui);
  }
  // ---- tree merge of the 8 warps' column states: 4..7 -> 0..3, 2..3 -> 0..1, 1 -> 0
#pragma unroll 1
  for (int span = 4; span >= 1; span >>= 1) {
    if (warp >= span && warp < 2 * span) {
      float2* dst = xch[warp - span];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float4* d4 = reinterpret_cast<float4*>(dst + 4 * lane + 128 * k);
        d4[0] = make_float4(cm[4 * k], cs[4 * k], cm[4 * k + 1], cs[4 * k + 1]);
        d4[1] = make_float4(cm[4 * k + 2], cs[4 * k + 2], cm[4 * k + 3], cs[4 * k + 3]);
      }
      if (lane == 0) dst[kOtFusedMaxM] = make_float2(bm, bs);
    }
    __syncthreads();
    if (warp < span) {
      const float2* src = xch[warp];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4* s4 = reinterpret_cast<const float4*>(src + 4 * lane + 128 * k);
        const float4 a = s4[0], bq = s4[1];
        lse_merge(cm[4 * k], cs[4 * k], a.x, a.y);
        lse_merge(cm[4 * k + 1], cs[4 * k + 1], a.z, a.w);
        lse_merge(cm[4 * k + 2], cs[4 * k + 2], bq.x, bq.y);
        lse_merge(cm[4 * k + 3], cs[4 * k + 3], bq.z, bq.w);
      }
      const float2 bb = src[kOtFusedMaxM];
      lse_merge(bm, bs, bb.x, bb.y);
    }
    __syncthreads();
  }
  if (warp == 0) {
    float2* prow = part_out + ((size_t)b * max_parts + blockIdx.x) * ld_part;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int j = 4 * lane + 128 * k;
      if (j < d.m) {     // ld_part is a multiple of 4 and > m, so the 4-wide store stays inside the row
        float4* dst = reinterpret_cast<float4*>(prow + j);
        dst[0] = make_float4(cm[4 * k], cs[4 * k], cm[4 * k + 1], cs[4 * k + 1]);
        dst[1] = make_float4(cm[4 * k + 2], cs[4 * k + 2], cm[4 * k + 3], cs[4 * k + 3]);
      }
    }
    __syncwarp();
    if (lane == 0) prow[d.m] = make_float2(bm, bs);
  }
}

// v from the partials of the last iteration
__global__ void __launch_bounds__(256) ot_finish_kernel(OtParams p, const float2* __restrict__ partials,
                                                        int max_parts, int ld_part) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * 256 + threadIdx.x;
  const PairDims d = pair_dims(p, b);
  if (d.n == 0 || d.m == 0 || j > d.m) return;
  const int parts = cdiv(d.n + 1, kOtRowsPerCta);
  const float2* col = partials + (size_t)b * max_parts * ld_part + j;
  float mm = kOtNegBig, ss = 0.f;
  for (int q0 = 0; q0 < parts; q0 += 18) {
    float2 t[18];
#pragma unroll
    for (int q = 0; q < 18; ++q)
      t[q] = q0 + q < parts ? __ldg(col + (size_t)(q0 + q) * ld_part) : make_float2(kOtNegBig, 0.f);
#pragma unroll
    for (int q = 0; q < 18; ++q) lse_merge(mm, ss, t[q].x, t[q].y);
  }
  p.v[(size_t)b * p.ld_uv + j] = (j == d.m ? d.nu_bin : d.norm) - (logf(ss) + mm);
}

bool ot_fused_supported(const OtParams& p) {
  return p.M <= kOtFusedMaxM && p.ldS % 4 == 0 && p.strideS % 4 == 0 && (reinterpret_cast<uintptr_t>(p.S) & 15) == 0;
}

// floats of scratch for `pairs` pairs: two ping-pong sets of partial rows
size_t ot_fused_scratch_floats(int pairs, int N, int M) {
  return (size_t)2 * pairs * ot_fused_parts(N) * round_up(M + 1, 4) * 2;
}

// `iters` full Sinkhorn iterations (u update then v update) starting from u = v = 0; leaves u and v in p.u / p.v
void launch_ot_sinkhorn_fused(LaunchCtx& ctx, const OtParams& p, int iters, float* scratch) {
  if (iters <= 0) return;
  const int max_parts = ot_fused_parts(p.N), ld_part = round_up(p.M + 1, 4);
  float2* part[2];
  part[0] = reinterpret_cast<float2*>(scratch);
  part[1] = part[0] + (size_t)p.B * max_parts * ld_part;
  for (int it = 0; it < iters; ++it) {
    ProfScope prof__(ctx, "ot_iter_fused");
    dim3 grid(max_parts, p.B);
    ot_iter_kernel<<<grid, 256, 0, ctx.stream>>>(p, part[(it + 1) & 1], part[it & 1], max_parts, ld_part, it == 0);
    B200M_LAUNCH_CHECK(ctx, "ot_iter_fused");
  }
  ProfScope prof__(ctx, "ot_finish");
  dim3 grid(cdiv(p.M + 1, 256), p.B);
  ot_finish_kernel<<<grid, 256, 0, ctx.stream>>>(p, part[(iters - 1) & 1], max_parts, ld_part);
  B200M_LAUNCH_CHECK(ctx, "ot_finish");
}

// Z = couplings + u + v - norm, dense (stage API; full sizes)
__global__ void ot_write_Z_kernel(OtParams p, float* __restrict__ Z) {
  const int b = blockIdx.z, i = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > p.M) return;
  const PairDims d = pair_dims(p, b);
  float c = (i == p.N || j == p.M) ? p.alpha : p.S[(size_t)b * p.strideS + (size_t)i * p.ldS + j];
  float z = ((c + p.u[(size_t)b * p.ld_uv + i]) + p.v[(size_t)b * p.ld_uv + j]) - d.norm;
  Z[((size_t)b * (p.N + 1) + i) * (p.M + 1) + j] = z;
}

void launch_ot_write_Z(LaunchCtx& ctx, const OtParams& p, float* Z) {
  dim3 grid(cdiv(p.M + 1, 256), p.N + 1, p.B);
  ot_write_Z_kernel<<<grid, 256, 0, ctx.stream>>>(p, Z);
  B200M_LAUNCH_CHECK(ctx, "ot_write_Z");
}

// ---- argmax over rows / columns of Z[:, :-1, :-1]  (scores.max(2), scores.max(1); first index on ties)
struct ZSource {
  const float* S; int ldS; long long strideS; const float* u; const float* v; int ld_uv;
  const float* Z; int ldZ; long long strideZ;   // dense alternative
};
__device__ __forceinline__ float z_at(const ZSource& z, int b, int i, int j, float norm) {
  if (z.Z) return z.Z[(size_t)b * z.strideZ + (size_t)i * z.ldZ + j];
  float c = z.S[(size_t)b * z.strideS + (size_t)i * z.ldS + j];
  return ((c + z.u[(size_t)b * z.ld_uv + i]) + z.v[(size_t)b * z.ld_uv + j]) - norm;
}

__global__ void __launch_bounds__(256) row_argmax_kernel(ZSource z, const int* counts0, const int* counts1,
                                                         int N, int M, int* __restrict__ idx0,
                                                         float* __restrict__ max0, int ld) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int n = counts0 ? counts0[b] : N, m = counts1 ? counts1[b] : M;
  if (i >= n || m == 0) return;
  const float norm = -logf((float)m + (float)n);
  float best = -INFINITY;
  int bj = 0x7fffffff;
  for (int j = lane; j < m; j += 32) {
    float val = z_at(z, b, i, j, norm);
    if (val > best) { best = val; bj = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oj = __shfl_xor_sync(0xffffffffu, bj, o);
    if (ov > best || (ov == best && oj < bj)) { best = ov; bj = oj; }
  }
  if (lane == 0) { idx0[(size_t)b * ld + i] = bj; max0[(size_t)b * ld + i] = best; }
}

__global__ void __launch_bounds__(256) col_argmax_kernel(ZSource z, const int* counts0, const int* counts1,
                                                         int N, int M, int* __restrict__ idx1, int ld) {
  __shared__ float rv[8][33];
  __shared__ int ri[8][33];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  const int n = counts0 ? counts0[b] : N, m = counts1 ? counts1[b] : M;
  if (n == 0 || blockIdx.x * 32 >= m) return;   // uniform per block
  const float norm = -logf((float)m + (float)n);
  float best = -INFINITY;
  int bi = 0x7fffffff;
  if (j < m)
    for (int i = w; i < n; i += 8) {
      float val = z_at(z, b, i, j, norm);
      if (val > best) { best = val; bi = i; }
    }
  rv[w][lane] = best;
  ri[w][lane] = bi;
  __syncthreads();
  if (w == 0 && j < m) {
    for (int k = 1; k < 8; ++k) {
      float ov = rv[k][lane];
      int oi = ri[k][lane];
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    idx1[(size_t)b * ld + j] = bi;
  }
}

static void launch_argmax_common(LaunchCtx& ctx, const ZSource& z, const int* c0, const int* c1, int B, int N,
                                 int M, int* idx0, float* max0, int* idx1, int ld) {
  ProfScope prof__(ctx, "argmax");
  if (N <= 0 || M <= 0) return;
  dim3 g0(cdiv(N, 8), B);
  row_argmax_kernel<<<g0, 256, 0, ctx.stream>>>(z, c0, c1, N, M, idx0, max0, ld);
  B200M_LAUNCH_CHECK(ctx, "row_argmax");
  dim3 g1(cdiv(M, 32), B);
  col_argmax_kernel<<<g1, 256, 0, ctx.stream>>>(z, c0, c1, N, M, idx1, ld);
  B200M_LAUNCH_CHECK(ctx, "col_argmax");
}

void launch_ot_argmax(LaunchCtx& ctx, const OtParams& p, int* idx0, float* max0, int* idx1) {
  ZSource z{p.S, p.ldS, p.strideS, p.u, p.v, p.ld_uv, nullptr, 0, 0};
  launch_argmax_common(ctx, z, p.counts0, p.counts1, p.B, p.N, p.M, idx0, max0, idx1, p.ld_uv);
}

void launch_dense_argmax(LaunchCtx& ctx, const float* Z, int B, int N, int M, int* idx0, float* max0, int* idx1,
                         int ld) {
  ZSource z{nullptr, 0, 0, nullptr, nullptr, 0, Z, M + 1, (long long)(N + 1) * (M + 1)};
  launch_argmax_common(ctx, z, nullptr, nullptr, B, N, M, idx0, max0, idx1, ld);
}

// mutual check + exp + threshold (:270-278); indices widened to int64 at the boundary
__global__ void match_select_kernel(const int* __restrict__ idx0, const float* __restrict__ max0,
                                    const int* __restrict__ idx1, int ld, const int* counts0, const int* counts1,
                                    int N, int M, float thr, long long* __restrict__ matches0,
                                    long long* __restrict__ matches1, float* __restrict__ ms0,
                                    float* __restrict__ ms1) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = counts0 ? counts0[b] : N, m = counts1 ? counts1[b] : M;
  const int* i0 = idx0 + (size_t)b * ld;
  const int* i1 = idx1 + (size_t)b * ld;
  const float* mx = max0 + (size_t)b * ld;
  const bool empty = (n == 0 || m == 0);
  if (t < N) {
    long long mt = -1;
    float sc = 0.f;
    if (!empty && t < n) {
      int j = i0[t];
      bool mutual = (i1[j] == t);
      sc = mutual ? expf(mx[t]) : 0.f;
      if (mutual && sc > thr) mt = j;
    }
    matches0[(size_t)b * N + t] = mt;
    ms0[(size_t)b * N + t] = sc;
  }
  if (t < M) {
    long long mt = -1;
    float sc = 0.f;
    if (!empty && t < m) {
      int i = i1[t];
      bool mutual1 = (i0[i] == t);
      // mscores0[i] and valid0[i] of the row this column points at
      bool mutual0 = (i1[i0[i]] == i);
      float s0 = mutual0 ? expf(mx[i]) : 0.f;
      sc = mutual1 ? s0 : 0.f;
      if (mutual1 && mutual0 && s0 > thr) mt = i;
    }
    matches1[(size_t)b * M + t] = mt;
    ms1[(size_t)b * M + t] = sc;
  }
}

void launch_match_select(LaunchCtx& ctx, const int* idx0, const float* max0, const int* idx1, int ld,
                         const int* counts0, const int* counts1, int B, int N, int M, float thr,
                         long long* matches0, long long* matches1, float* ms0, float* ms1) {
  ProfScope prof__(ctx, "match_select");
  int T = N > M ? N : M;
  if (T <= 0) return;
  dim3 grid(cdiv(T, 256), B);
  match_select_kernel<<<grid, 256, 0, ctx.stream>>>(idx0, max0, idx1, ld, counts0, counts1, N, M, thr, matches0,
                                                    matches1, ms0, ms1);
  B200M_LAUNCH_CHECK(ctx, "match_select");
}

}  // namespace b200m
