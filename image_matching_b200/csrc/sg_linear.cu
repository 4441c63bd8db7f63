// SuperGlue linear layers (Conv1d k=1 == per-token GEMM) and layout helpers, fp32 CUDA-core path.
// Reference: superglue/models/superglue_test.py:49-60 (MLP), :73-82 (KeypointEncoder), :98-107 (proj/merge),
// :110-119 (AttentionalPropagation.mlp), :256-260 (final_proj + score einsum).
#include <cuda_fp16.h>
#include "kernels.cuh"

namespace b200m {

// C[M,N] (+)= alpha * A[M,K] * Bw[N,K]^T + bias, optional ReLU.  Both operands K-contiguous.
// Block 256 threads computes a 128x64 tile; thread = 8 rows x 4 cols; K chunks of 16 staged through
// registers into transposed shared tiles (conflict-free reads), register double buffering.
constexpr int GM = 128, GN = 64, GK = 16;

__global__ void __launch_bounds__(256) gemm_tn_kernel(GemmParams p) {
  pdl_trigger();   // PDL (common.cuh): the next kernel may be scheduled; nothing is read or written before the wait
  pdl_wait();
  __shared__ float As[2][GK][GM + 4];
  __shared__ float Bs[2][GK][GN + 4];
  const int tid = threadIdx.x;
  const int bz = blockIdx.z;
  const float* A = p.A + (size_t)bz * p.strideA;
  const float* Bw = p.Bw + (size_t)bz * p.strideB;
  float* C = p.C + (size_t)bz * p.strideC;
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  const int tx = tid & 15, ty = tid >> 4;   // cols tx*4.., rows ty*8..
  // loader mapping: A tile = 128 rows x 4 float4 -> 512 float4, two per thread; B tile = 64 x 4 -> one
  const int a_row0 = tid >> 2, a_k4 = tid & 3;          // rows a_row0 and a_row0 + 64
  const int b_row = tid >> 2, b_k4 = tid & 3;
  float4 ra[2], rb;
  auto gload = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int r = m0 + a_row0 + h * 64, k = k0 + a_k4 * 4;
      ra[h] = (r < p.M && k < p.K) ? *reinterpret_cast<const float4*>(A + (size_t)r * p.lda + k)
                                   : make_float4(0, 0, 0, 0);
    }
    int r = n0 + b_row, k = k0 + b_k4 * 4;
    rb = (r < p.N && k < p.K) ? *reinterpret_cast<const float4*>(Bw + (size_t)r * p.ldb + k)
                              : make_float4(0, 0, 0, 0);
  };
  auto sstore = [&](int s) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int r = a_row0 + h * 64;
      As[s][a_k4 * 4 + 0][r] = ra[h].x; As[s][a_k4 * 4 + 1][r] = ra[h].y;
      As[s][a_k4 * 4 + 2][r] = ra[h].z; As[s][a_k4 * 4 + 3][r] = ra[h].w;
    }
    Bs[s][b_k4 * 4 + 0][b_row] = rb.x; Bs[s][b_k4 * 4 + 1][b_row] = rb.y;
    Bs[s][b_k4 * 4 + 2][b_row] = rb.z; Bs[s][b_k4 * 4 + 3][b_row] = rb.w;
  };
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = cdiv(p.K, GK);
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int s = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * GK);
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[s][k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[s][k][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[s][k][tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(s ^ 1);
      __syncthreads();
    }
  }
  // epilogue
  float bz4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = n0 + tx * 4 + j;
    if (p.bias && c < p.N) bz4[j] = p.bias[c];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r = m0 + ty * 8 + i;
    if (r >= p.M) continue;
    float* crow = C + (size_t)r * p.ldc;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = n0 + tx * 4 + j;
      if (c >= p.N) continue;
      float v = p.alpha * acc[i][j] + bz4[j];
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.accumulate) v += crow[c];
      if (p.out_f16) {
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        const size_t o = (size_t)bz * p.strideC + (size_t)r * p.ldc + c;
        reinterpret_cast<__half*>(p.C)[o] = hi;
        reinterpret_cast<__half*>(p.C_lo)[o] = lo;
        if (p.VT && c >= p.vt_col0) {
          const int blk = r / p.vt_np, rr = r - blk * p.vt_np;
          const size_t ov = ((size_t)blk * (p.N - p.vt_col0) + (c - p.vt_col0)) * p.vt_np + rr;
          reinterpret_cast<__half*>(p.VT)[ov] = hi;
          reinterpret_cast<__half*>(p.VT_lo)[ov] = lo;
        }
      } else {
        crow[c] = v;
      }
    }
  }
}

void launch_gemm(LaunchCtx& ctx, const GemmParams& p) {
  ProfScope prof__(ctx, "gemm");
  if (p.M <= 0 || p.N <= 0 || p.batch <= 0) return;
  dim3 grid(cdiv(p.N, GN), cdiv(p.M, GM), p.batch);
  launch_pdl(ctx, kPdlGemm, gemm_tn_kernel, dim3(grid), dim3(256), 0, p);
  B200M_LAUNCH_CHECK(ctx, "gemm_tn");
}

// ------------------------------------------------------------------------------------------------
// (B,C,N) channel-major  ->  token-major rows [B][Np][ld] (columns [0,C)); rows >= N are zeroed.
// 32x32 shared-memory transpose so both sides are coalesced.
__global__ void bcn_to_tokens_kernel(const float* __restrict__ in, int C, int N, float* __restrict__ out,
                                     int Np, int ld) {
  __shared__ float t[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const float* src = in + (size_t)b * C * N;
  float* dst = out + (size_t)b * Np * ld;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, nn = n0 + threadIdx.x;
    t[i][threadIdx.x] = (c < C && nn < N) ? src[(size_t)c * N + nn] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int nn = n0 + i, c = c0 + threadIdx.x;
    if (nn < Np && c < C) dst[(size_t)nn * ld + c] = t[threadIdx.x][i];
  }
}

void launch_bcn_to_tokens(LaunchCtx& ctx, const float* in, int B, int C, int N, float* out, int Np, int ld) {
  ProfScope prof__(ctx, "bcn_to_tokens");
  dim3 grid(cdiv(Np, 32), cdiv(C, 32), B), block(32, 8);
  bcn_to_tokens_kernel<<<grid, block, 0, ctx.stream>>>(in, C, N, out, Np, ld);
  B200M_LAUNCH_CHECK(ctx, "bcn_to_tokens");
}

__global__ void tokens_to_bcn_kernel(const float* __restrict__ in, int Np, int ld, float* __restrict__ out,
                                     int C, int N) {
  __shared__ float t[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const float* src = in + (size_t)b * Np * ld;
  float* dst = out + (size_t)b * C * N;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int nn = n0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (nn < N && c < C) ? src[(size_t)nn * ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, nn = n0 + threadIdx.x;
    if (c < C && nn < N) dst[(size_t)c * N + nn] = t[threadIdx.x][i];
  }
}

void launch_tokens_to_bcn(LaunchCtx& ctx, const float* in, int Np, int ld, float* out, int B, int C, int N) {
  ProfScope prof__(ctx, "tokens_to_bcn");
  dim3 grid(cdiv(N, 32), cdiv(C, 32), B), block(32, 8);
  tokens_to_bcn_kernel<<<grid, block, 0, ctx.stream>>>(in, Np, ld, out, C, N);
  B200M_LAUNCH_CHECK(ctx, "tokens_to_bcn");
}

// fp32 q|k|v rows -> fp16 hi/lo planes + transposed V planes (test hook for the attention kernel)
__global__ void qkv_to_f16_planes_kernel(const float* __restrict__ qkv, __half* __restrict__ hi, __half* __restrict__ lo,
                                         __half* __restrict__ vh, __half* __restrict__ vl, int Np, int D) {
  const int blk = blockIdx.z, row = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= 3 * D) return;
  const size_t o = ((size_t)blk * Np + row) * 3 * D + c;
  const float v = qkv[o];
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn(v - __half2float(h));
  hi[o] = h;
  lo[o] = l;
  if (c >= 2 * D) {
    const size_t ov = ((size_t)blk * D + (c - 2 * D)) * Np + row;
    vh[ov] = h;
    vl[ov] = l;
  }
}

void launch_qkv_to_f16_planes(LaunchCtx& ctx, const float* qkv, void* hi, void* lo, void* vt_hi, void* vt_lo,
                              int blocks, int Np, int D) {
  ProfScope prof__(ctx, "qkv_to_f16_planes");
  dim3 grid(cdiv(3 * D, 128), Np, blocks);
  qkv_to_f16_planes_kernel<<<grid, 128, 0, ctx.stream>>>(qkv, reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo),
                                                         reinterpret_cast<__half*>(vt_hi), reinterpret_cast<__half*>(vt_lo), Np, D);
  B200M_LAUNCH_CHECK(ctx, "qkv_to_f16_planes");
}

// fp32 columns [0, cols) of a row-major matrix -> fp16 operand planes hi = fp16(v), lo = fp16((v - hi) * 2048): the
// A-operand format tc_gemm.cu consumes without its in-kernel split (the D != 128 GNN layers, api.cu sg_gnn)
__global__ void split_planes_kernel(const float* __restrict__ in, int ld_in, __half* __restrict__ hi,
                                    __half* __restrict__ lo, int ld_out, size_t rows, int cols8) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * (size_t)cols8) return;
  const size_t r = i / cols8;
  const int c = (int)(i - r * cols8) * 8;
  const float4 a = *reinterpret_cast<const float4*>(in + r * ld_in + c);
  const float4 b = *reinterpret_cast<const float4*>(in + r * ld_in + c + 4);
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  __align__(16) __half h[8], l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    h[j] = __float2half_rn(v[j]);
    l[j] = __float2half_rn((v[j] - __half2float(h[j])) * 2048.f);
  }
  *reinterpret_cast<uint4*>(hi + r * ld_out + c) = *reinterpret_cast<const uint4*>(h);
  *reinterpret_cast<uint4*>(lo + r * ld_out + c) = *reinterpret_cast<const uint4*>(l);
}

void launch_split_planes(LaunchCtx& ctx, const float* in, int ld_in, void* hi, void* lo, int ld_out, size_t rows,
                         int cols) {
  ProfScope prof__(ctx, "split_planes");
  const size_t n = rows * (size_t)(cols / 8);
  if (n == 0) return;
  launch_pdl(ctx, kPdlGemm, split_planes_kernel, dim3((unsigned)cdivz(n, 256)), dim3(256), 0, in, ld_in,
             reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo), ld_out, rows, cols / 8);
  B200M_LAUNCH_CHECK(ctx, "split_planes");
}

// normalize_keypoints (:63-70) fused with the cat([kpts^T, scores]) of KeypointEncoder.forward (:80-82):
// rows of (x_norm, y_norm, score, 0) -- K padded 3 -> 4 so the first layer is a float4 GEMM.
__global__ void kenc_input_kernel(const float* __restrict__ kpts, const float* __restrict__ scores, int N, int Np,
                                  float cx, float cy, float scale, float4* __restrict__ out, int total) {
  pdl_trigger();   // PDL (common.cuh): the next kernel may be scheduled; nothing is read or written before the wait
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int b = i / Np, k = i - b * Np;
  float4 v = make_float4(0, 0, 0, 0);
  if (k < N) {
    size_t s = (size_t)b * N + k;
    v.x = (kpts[2 * s] - cx) / scale;
    v.y = (kpts[2 * s + 1] - cy) / scale;
    v.z = scores[s];
  }
  out[i] = v;
}

void launch_kenc_input(LaunchCtx& ctx, const float* kpts, const float* scores, int B, int N, int Np,
                       float cx, float cy, float scale, float* out4) {
  ProfScope prof__(ctx, "kenc_input");
  int total = B * Np;
  launch_pdl(ctx, kPdlGemm, kenc_input_kernel, dim3(cdiv(total, 256)), dim3(256), 0, kpts, scores, N, Np, cx, cy, scale,
                                                                reinterpret_cast<float4*>(out4), total);
  B200M_LAUNCH_CHECK(ctx, "kenc_input");
}

}  // namespace b200m
