// SuperPoint encoder / head convolutions, fp32 CUDA-core path (exact-fp32 accumulate).
// Reference: superpoint/models/unet_parts.py:10-48, superpoint/models/superpoint_test.py:113-126.
// BatchNorm is folded into (w, b) at pack time; ReLU and the 2x2 max-pool are fused in the epilogue.
#include <cuda_fp16.h>
#include <algorithm>
#include "kernels.cuh"

namespace b200m {

// ------------------------------------------------------------------------------------------------
// conv1: 1 -> 64 channels, 3x3, zero pad, +bias, ReLU.  A stencil, not a GEMM (K = 9).
// Block (32,8) = 32x8 pixels; every thread produces the 64 output channels of one pixel and stores
// them as 16 float4 (one per channel group): a warp writes 512 contiguous bytes per group.
// HBM-bound on the 64-channel fp32 output (256 B per pixel).
// round-to-nearest onto the tf32 grid (10 explicit mantissa bits): unbiased, unlike truncation
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }

__global__ void __launch_bounds__(256) conv1_direct_kernel(const float* __restrict__ img,
                                                           const float* __restrict__ w9x64,
                                                           const float* __restrict__ bias,
                                                           float* __restrict__ out,
                                                           float* __restrict__ out_lo, int H, int W) {
  __shared__ float4 sw[9 * 16];
  __shared__ float4 sb[16];
  __shared__ float tile[10][34];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int n = blockIdx.z;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
  if (tid < 144) sw[tid] = reinterpret_cast<const float4*>(w9x64)[tid];
  if (tid < 16) sb[tid] = reinterpret_cast<const float4*>(bias)[tid];
  const float* im = img + (size_t)n * H * W;
  for (int i = tid; i < 10 * 34; i += 256) {
    int r = i / 34, c = i % 34;
    int gy = y0 - 1 + r, gx = x0 - 1 + c;
    tile[r][c] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? im[(size_t)gy * W + gx] : 0.f;
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= W || y >= H) return;
  float v[9];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) v[ky * 3 + kx] = tile[threadIdx.y + ky][threadIdx.x + kx];
  const size_t plane = (size_t)H * W;
  float4* o = reinterpret_cast<float4*>(out) + (size_t)n * 16 * plane + (size_t)y * W + x;
  uint4* oh = reinterpret_cast<uint4*>(out) + (size_t)n * 8 * plane + (size_t)y * W + x;      // fp16 C8-planar
  uint4* ol = reinterpret_cast<uint4*>(out_lo) + (size_t)n * 8 * plane + (size_t)y * W + x;
  float prev[4];
#pragma unroll 4
  for (int g = 0; g < 16; ++g) {
    float4 a = sb[g];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float4 wv = sw[t * 16 + g];
      a.x = fmaf(v[t], wv.x, a.x);
      a.y = fmaf(v[t], wv.y, a.y);
      a.z = fmaf(v[t], wv.z, a.z);
      a.w = fmaf(v[t], wv.w, a.w);
    }
    a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
    if (out_lo) {
      if ((g & 1) == 0) {
        prev[0] = a.x; prev[1] = a.y; prev[2] = a.z; prev[3] = a.w;
      } else {
        const float e[8] = {prev[0], prev[1], prev[2], prev[3], a.x, a.y, a.z, a.w};
        __half2 h2[4], l2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __half ha = __float2half_rn(e[2 * j]), hb = __float2half_rn(e[2 * j + 1]);
          h2[j] = __halves2half2(ha, hb);
          l2[j] = __halves2half2(__float2half_rn((e[2 * j] - __half2float(ha)) * 2048.f),
                                 __float2half_rn((e[2 * j + 1] - __half2float(hb)) * 2048.f));
        }
        oh[(size_t)(g >> 1) * plane] = *reinterpret_cast<uint4*>(h2);
        ol[(size_t)(g >> 1) * plane] = *reinterpret_cast<uint4*>(l2);
      }
    } else {
      o[(size_t)g * plane] = a;
    }
  }
}

// datasets/SSHIDataset.py:26-29 (`img / 255.` in float64, `.float()` by the caller) on the device, so a caller can
// upload the 8-bit image (4x fewer PCIe bytes than fp32)
__global__ void u8_to_unit_f32_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = i; k < n; k += stride) out[k] = __fdiv_rn((float)in[k], 255.f);
}

void launch_u8_to_unit_f32(LaunchCtx& ctx, const uint8_t* in, float* out, size_t n) {
  ProfScope prof__(ctx, "u8_to_f32");
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
  u8_to_unit_f32_kernel<<<blocks, 256, 0, ctx.stream>>>(in, out, n);
  B200M_LAUNCH_CHECK(ctx, "u8_to_f32");
}

void launch_conv1_direct(LaunchCtx& ctx, const float* img, const float* w9x64, const float* bias,
                         float* out, float* out_lo, int n, int H, int W) {
  ProfScope prof__(ctx, "conv1_direct");
  dim3 grid(cdiv(W, 32), cdiv(H, 8), n), block(32, 8);
  conv1_direct_kernel<<<grid, block, 0, ctx.stream>>>(img, w9x64, bias, out, out_lo, H, W);
  B200M_LAUNCH_CHECK(ctx, "conv1_direct");
}

// ------------------------------------------------------------------------------------------------
// Generic KSxKS (3 or 1) convolution on C4-planar activations, fp32 FFMA.
//   block = 256 threads, output tile = 16x16 pixels x 64 output channels
//   thread = (4 wide x 2 tall) pixels x 8 output channels  (64 accumulators)
//   K loop: chunks of 8 input channels; per chunk the (16+2)^2 halo tile (2 float4 planes) and the
//   [tap][8][64] weight slab are staged with cp.async, double buffered.
template <int KS>
struct ConvSmem {
  static constexpr int HALO = KS / 2;
  static constexpr int TW = 16 + 2 * HALO;
  static constexpr int TAPS = KS * KS;
  static constexpr int IN_F4 = 2 * TW * TW;          // float4 per stage
  static constexpr int W_F4 = TAPS * 8 * 64 / 4;     // float4 per stage
  static constexpr int STAGE_F4 = IN_F4 + W_F4;
  static constexpr size_t BYTES = 2 * (size_t)STAGE_F4 * sizeof(float4);
};

__device__ __forceinline__ float f4c(const float4& v, int i) {
  return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

template <int KS, bool POOL>
__global__ void __launch_bounds__(256, 2) conv_c4_kernel(ConvParams p) {
  using SM = ConvSmem<KS>;
  constexpr int HALO = SM::HALO, TW = SM::TW;
  extern __shared__ float4 smem[];
  const int tid = threadIdx.x;
  const int cg = tid & 7;          // 8 output channels cg*8..cg*8+7 (within the 64-channel block)
  const int pg = tid >> 3;         // pixel group 0..31
  const int pgx = pg & 3, pgy = pg >> 2;
  const int tiles_x = cdiv(p.W, 16);
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int x0 = tx * 16, y0 = ty * 16;
  const int cb = blockIdx.y, n = blockIdx.z;
  const int nchunks = p.cin / 8;
  const size_t plane = (size_t)p.H * p.W;
  const float4* in4 = reinterpret_cast<const float4*>(p.in) + ((size_t)n * p.in_c4_total + p.in_c4_off) * plane;
  const float4* w4 = reinterpret_cast<const float4*>(p.wpk) + (size_t)cb * nchunks * SM::W_F4;

  auto load_stage = [&](int cc, int s) {
    float4* sin = smem + (size_t)s * SM::STAGE_F4;
    float4* sw = sin + SM::IN_F4;
    for (int i = tid; i < SM::IN_F4; i += 256) {
      int c4 = i / (TW * TW);
      int rem = i - c4 * (TW * TW);
      int r = rem / TW, c = rem - r * TW;
      int gy = y0 - HALO + r, gx = x0 - HALO + c;
      bool ok = (gy >= 0) && (gy < p.H) && (gx >= 0) && (gx < p.W);
      const float4* src = ok ? in4 + (size_t)(cc * 2 + c4) * plane + (size_t)gy * p.W + gx : in4;
      cp_async16(sin + i, src, ok);
    }
    const float4* wsrc = w4 + (size_t)cc * SM::W_F4;
    for (int i = tid; i < SM::W_F4; i += 256) cp_async16(sw + i, wsrc + i, true);
    cp_async_commit();
  };

  float acc[2][4][8];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[a][b][c] = 0.f;

  load_stage(0, 0);
  for (int cc = 0; cc < nchunks; ++cc) {
    if (cc + 1 < nchunks) {
      load_stage(cc + 1, (cc + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float4* sin = smem + (size_t)(cc & 1) * SM::STAGE_F4;
    const float4* sw = sin + SM::IN_F4;
#pragma unroll 1
    for (int c4 = 0; c4 < 2; ++c4) {
#pragma unroll 1
      for (int r = 0; r < 2 + 2 * HALO; ++r) {
        float4 px[4 + 2 * HALO];
        const float4* rowp = sin + (c4 * TW + pgy * 2 + r) * TW + pgx * 4;
#pragma unroll
        for (int j = 0; j < 4 + 2 * HALO; ++j) px[j] = rowp[j];
#pragma unroll
        for (int oy = 0; oy < 2; ++oy) {
          const int ky = r - oy;
          if (ky < 0 || ky >= KS) continue;
#pragma unroll
          for (int kx = 0; kx < KS; ++kx) {
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) {
              const float4* wp = sw + (((ky * KS + kx) * 8 + c4 * 4 + ci) * 64 + cg * 8) / 4;
              const float4 wa = wp[0], wb = wp[1];
#pragma unroll
              for (int ox = 0; ox < 4; ++ox) {
                const float v = f4c(px[ox + kx], ci);
                float* a = acc[oy][ox];
                a[0] = fmaf(v, wa.x, a[0]); a[1] = fmaf(v, wa.y, a[1]);
                a[2] = fmaf(v, wa.z, a[2]); a[3] = fmaf(v, wa.w, a[3]);
                a[4] = fmaf(v, wb.x, a[4]); a[5] = fmaf(v, wb.y, a[5]);
                a[6] = fmaf(v, wb.z, a[6]); a[7] = fmaf(v, wb.w, a[7]);
              }
            }
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue: bias, ReLU, optional 2x2 max-pool, C4-planar store
  const int co = cb * 64 + cg * 8;
  float bz[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bz[j] = p.bias[co + j];
#pragma unroll
  for (int oy = 0; oy < 2; ++oy)
#pragma unroll
    for (int ox = 0; ox < 4; ++ox)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = acc[oy][ox][j] + bz[j];
        acc[oy][ox][j] = p.relu ? fmaxf(v, 0.f) : v;
      }
  const int g0 = p.out_c4_off + cb * 16 + cg * 2;
  if (POOL) {
    const int Ho = p.H / 2, Wo = p.W / 2;
    const int Y = (y0 + pgy * 2) >> 1;
    float4* o = reinterpret_cast<float4*>(p.out) + (size_t)n * p.out_c4_total * Ho * Wo;
#pragma unroll
    for (int oxp = 0; oxp < 2; ++oxp) {
      const int X = ((x0 + pgx * 4) >> 1) + oxp;
      if (Y < Ho && X < Wo) {
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          m[j] = fmaxf(fmaxf(acc[0][2 * oxp][j], acc[0][2 * oxp + 1][j]),
                       fmaxf(acc[1][2 * oxp][j], acc[1][2 * oxp + 1][j]));
        size_t off = (size_t)Y * Wo + X;
        o[(size_t)g0 * Ho * Wo + off] = make_float4(m[0], m[1], m[2], m[3]);
        o[(size_t)(g0 + 1) * Ho * Wo + off] = make_float4(m[4], m[5], m[6], m[7]);
      }
    }
  } else {
    float4* o = reinterpret_cast<float4*>(p.out) + (size_t)n * p.out_c4_total * plane;
#pragma unroll
    for (int oy = 0; oy < 2; ++oy) {
      const int y = y0 + pgy * 2 + oy;
#pragma unroll
      for (int ox = 0; ox < 4; ++ox) {
        const int x = x0 + pgx * 4 + ox;
        if (y < p.H && x < p.W) {
          const float* a = acc[oy][ox];
          size_t off = (size_t)y * p.W + x;
          o[(size_t)g0 * plane + off] = make_float4(a[0], a[1], a[2], a[3]);
          o[(size_t)(g0 + 1) * plane + off] = make_float4(a[4], a[5], a[6], a[7]);
        }
      }
    }
  }
}

template <int KS, bool POOL>
static void launch_conv_t(LaunchCtx& ctx, const ConvParams& p) {
  ProfScope prof__(ctx, KS == 3 ? "conv3x3_c4" : "conv1x1_c4");
  static SmemOptIn opt;
  auto kern = conv_c4_kernel<KS, POOL>;
  opt.ensure(kern, (int)ConvSmem<KS>::BYTES);
  dim3 grid(cdiv(p.W, 16) * cdiv(p.H, 16), p.cout_pad / 64, p.n);
  kern<<<grid, 256, ConvSmem<KS>::BYTES, ctx.stream>>>(p);
  B200M_LAUNCH_CHECK(ctx, "conv_c4");
}

void launch_conv(LaunchCtx& ctx, const ConvParams& p, int ksize, bool pool) {
  if (ksize == 3) {
    if (pool) launch_conv_t<3, true>(ctx, p); else launch_conv_t<3, false>(ctx, p);
  } else {
    if (pool) launch_conv_t<1, true>(ctx, p); else launch_conv_t<1, false>(ctx, p);
  }
}

// ------------------------------------------------------------------------------------------------
// Layout conversions at the stage-API boundary.
__global__ void c4_to_nchw_kernel(const float4* __restrict__ in, const float4* __restrict__ in_lo, int c4_total,
                                  int c4_off, int C, float* __restrict__ out, int HW, int normalize) {
  const int n = blockIdx.y;
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= HW) return;
  const float4* src = in + ((size_t)n * c4_total + c4_off) * HW + px;
  const float4* slo = in_lo ? in_lo + ((size_t)n * c4_total + c4_off) * HW + px : nullptr;
  float* dst = out + (size_t)n * C * HW + px;
  const int G = cdiv(C, 4);
  float nrm = 1.f;
  if (normalize) {
    float ss = 0.f;
    for (int g = 0; g < G; ++g) {
      float4 v = src[(size_t)g * HW];
      float e[4] = {v.x, v.y, v.z, v.w};
      for (int j = 0; j < 4; ++j)
        if (g * 4 + j < C) ss += e[j] * e[j];
    }
    nrm = sqrtf(ss);
  }
  for (int g = 0; g < G; ++g) {
    float4 v = src[(size_t)g * HW];
    if (slo) {
      float4 l = slo[(size_t)g * HW];
      v.x += l.x; v.y += l.y; v.z += l.z; v.w += l.w;
    }
    float e[4] = {v.x, v.y, v.z, v.w};
    for (int j = 0; j < 4; ++j) {
      int c = g * 4 + j;
      if (c < C) dst[(size_t)c * HW] = normalize ? e[j] / nrm : e[j];
    }
  }
}

__global__ void c4_split_kernel(const float4* __restrict__ in, float4* __restrict__ hi, float4* __restrict__ lo, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a = in[i], h, l;
  h.x = tf32_hi(a.x); h.y = tf32_hi(a.y); h.z = tf32_hi(a.z); h.w = tf32_hi(a.w);
  l.x = tf32_hi(a.x - h.x); l.y = tf32_hi(a.y - h.y); l.z = tf32_hi(a.z - h.z); l.w = tf32_hi(a.w - h.w);
  hi[i] = h;
  lo[i] = l;
}

void launch_c4_split(LaunchCtx& ctx, const float* in, float* hi, float* lo, size_t n_float4) {
  ProfScope prof__(ctx, "c4_split");
  c4_split_kernel<<<(unsigned)cdivz(n_float4, 256), 256, 0, ctx.stream>>>(
      reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(hi), reinterpret_cast<float4*>(lo), n_float4);
  B200M_LAUNCH_CHECK(ctx, "c4_split");
}

void launch_c4_to_nchw(LaunchCtx& ctx, const float* in, int c4_total, int c4_off, int C, float* out,
                       int n, int H, int W, bool l2_normalize, const float* in_lo) {
  ProfScope prof__(ctx, "c4_to_nchw");
  int HW = H * W;
  dim3 grid(cdiv(HW, 128), n);
  c4_to_nchw_kernel<<<grid, 128, 0, ctx.stream>>>(reinterpret_cast<const float4*>(in),
                                                   reinterpret_cast<const float4*>(in_lo), c4_total, c4_off, C,
                                                   out, HW, l2_normalize ? 1 : 0);
  B200M_LAUNCH_CHECK(ctx, "c4_to_nchw");
}

__global__ void nchw_to_c4_kernel(const float* __restrict__ in, int C, float4* __restrict__ out,
                                  int c4_total, int HW) {
  const int n = blockIdx.y;
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= HW) return;
  const float* src = in + (size_t)n * C * HW + px;
  float4* dst = out + (size_t)n * c4_total * HW + px;
  for (int g = 0; g < c4_total; ++g) {
    float e[4];
    for (int j = 0; j < 4; ++j) {
      int c = g * 4 + j;
      e[j] = c < C ? src[(size_t)c * HW] : 0.f;
    }
    dst[(size_t)g * HW] = make_float4(e[0], e[1], e[2], e[3]);
  }
}

void launch_nchw_to_c4(LaunchCtx& ctx, const float* in, int C, float* out, int c4_total, int n, int H, int W) {
  ProfScope prof__(ctx, "nchw_to_c4");
  int HW = H * W;
  dim3 grid(cdiv(HW, 128), n);
  nchw_to_c4_kernel<<<grid, 128, 0, ctx.stream>>>(in, C, reinterpret_cast<float4*>(out), c4_total, HW);
  B200M_LAUNCH_CHECK(ctx, "nchw_to_c4");
}

__global__ void nchw_to_c8_split_kernel(const float* __restrict__ in, int C, uint4* __restrict__ hi, uint4* __restrict__ lo,
                                        int HW) {
  const int n = blockIdx.y;
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= HW) return;
  const float* src = in + (size_t)n * C * HW + px;
  const int U = C / 8;
  for (int u = 0; u < U; ++u) {
    __half2 h2[4], l2[4];
    for (int j = 0; j < 4; ++j) {
      const float a = src[(size_t)(u * 8 + 2 * j) * HW], b = src[(size_t)(u * 8 + 2 * j + 1) * HW];
      const __half ha = __float2half_rn(a), hb = __float2half_rn(b);
      h2[j] = __halves2half2(ha, hb);
      l2[j] = __halves2half2(__float2half_rn((a - __half2float(ha)) * 2048.f),
                             __float2half_rn((b - __half2float(hb)) * 2048.f));
    }
    hi[((size_t)n * U + u) * HW + px] = *reinterpret_cast<uint4*>(h2);
    lo[((size_t)n * U + u) * HW + px] = *reinterpret_cast<uint4*>(l2);
  }
}

void launch_nchw_to_c8_split(LaunchCtx& ctx, const float* in, int C, void* hi, void* lo, int n, int H, int W) {
  ProfScope prof__(ctx, "nchw_to_c8_split");
  dim3 grid(cdiv(H * W, 128), n);
  nchw_to_c8_split_kernel<<<grid, 128, 0, ctx.stream>>>(in, C, reinterpret_cast<uint4*>(hi), reinterpret_cast<uint4*>(lo), H * W);
  B200M_LAUNCH_CHECK(ctx, "nchw_to_c8_split");
}

__global__ void c8_to_nchw_kernel(const uint4* __restrict__ hi, const uint4* __restrict__ lo, int c8_total, int C,
                                  float* __restrict__ out, int HW) {
  const int n = blockIdx.y;
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= HW) return;
  for (int u = 0; u < cdiv(C, 8); ++u) {
    uint4 h = hi[((size_t)n * c8_total + u) * HW + px], l = lo[((size_t)n * c8_total + u) * HW + px];
    const __half* hh = reinterpret_cast<const __half*>(&h);
    const __half* ll = reinterpret_cast<const __half*>(&l);
    for (int j = 0; j < 8; ++j)
      if (u * 8 + j < C)
        out[((size_t)n * C + u * 8 + j) * HW + px] = __half2float(hh[j]) + __half2float(ll[j]) * (1.f / 2048.f);
  }
}

void launch_c8_to_nchw(LaunchCtx& ctx, const void* hi, const void* lo, int c8_total, int C, float* out, int n, int H, int W) {
  ProfScope prof__(ctx, "c8_to_nchw");
  dim3 grid(cdiv(H * W, 128), n);
  c8_to_nchw_kernel<<<grid, 128, 0, ctx.stream>>>(reinterpret_cast<const uint4*>(hi), reinterpret_cast<const uint4*>(lo),
                                                   c8_total, C, out, H * W);
  B200M_LAUNCH_CHECK(ctx, "c8_to_nchw");
}

// desc / ||desc||_2 over channels, no eps (superpoint_test.py:125-126)
__global__ void c4_sumsq_kernel(const float4* __restrict__ in, int in_c4_total, int in_c4_off,
                                float* __restrict__ sumsq, int G, int HW) {
  const int n = blockIdx.y;
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= HW) return;
  const float4* src = in + ((size_t)n * in_c4_total + in_c4_off) * HW + px;
  float ss = 0.f;
  for (int g = 0; g < G; ++g) {
    float4 v = src[(size_t)g * HW];
    ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
  }
  sumsq[(size_t)n * HW + px] = ss;
}

void launch_c4_sumsq(LaunchCtx& ctx, const float* in, int in_c4_total, int in_c4_off, float* sumsq, int C, int n, int H,
                     int W) {
  ProfScope prof__(ctx, "c4_sumsq");
  int HW = H * W;
  dim3 grid(cdiv(HW, 128), n);
  c4_sumsq_kernel<<<grid, 128, 0, ctx.stream>>>(reinterpret_cast<const float4*>(in), in_c4_total, in_c4_off, sumsq,
                                                C / 4, HW);
  B200M_LAUNCH_CHECK(ctx, "c4_sumsq");
}

}  // namespace b200m
