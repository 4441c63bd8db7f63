// SuperPoint detector post-processing and descriptor sampling (HBM / shared-memory bound, fp32,
// compare-only NMS => bit-exact given the same heat-map).
// Reference: superpoint/models/superpoint_test.py:7-52 and :128-155.
#include "kernels.cuh"

namespace b200m {

// ------------------------------------------------------------------------------------------------
// softmax over the 65 detector channels, drop the dustbin, depth-to-space x8  (:128-131)
// One thread per coarse cell; a warp covers 32 consecutive cells of a row, so each of the 8 output
// rows receives one 1024 B contiguous run per warp.
__global__ void __launch_bounds__(128) softmax_heat_kernel(const float4* __restrict__ semi, int c4_total,
                                                           float* __restrict__ heat, int hc, int wc) {
  const int n = blockIdx.z;
  const int cx = blockIdx.x * blockDim.x + threadIdx.x;
  const int cy = blockIdx.y;
  if (cx >= wc) return;
  const size_t plane = (size_t)hc * wc;
  const float4* src = semi + (size_t)n * c4_total * plane + (size_t)cy * wc + cx;
  float v[68];
#pragma unroll
  for (int g = 0; g < 17; ++g) {
    float4 t = src[(size_t)g * plane];
    v[4 * g] = t.x; v[4 * g + 1] = t.y; v[4 * g + 2] = t.z; v[4 * g + 3] = t.w;
  }
  float m = v[0];
#pragma unroll
  for (int c = 1; c < 65; ++c) m = fmaxf(m, v[c]);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 65; ++c) {
    v[c] = expf(v[c] - m);
    s += v[c];
  }
  const int W8 = wc * 8;
  float* dst = heat + (size_t)n * (hc * 8) * W8 + (size_t)(cy * 8) * W8 + cx * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 a = make_float4(v[8 * i] / s, v[8 * i + 1] / s, v[8 * i + 2] / s, v[8 * i + 3] / s);
    float4 b = make_float4(v[8 * i + 4] / s, v[8 * i + 5] / s, v[8 * i + 6] / s, v[8 * i + 7] / s);
    float4* d4 = reinterpret_cast<float4*>(dst + (size_t)i * W8);
    d4[0] = a;
    d4[1] = b;
  }
}

void launch_softmax_heat(LaunchCtx& ctx, const float* semi_c4, int c4_total, float* heat, int n, int hc, int wc) {
  ProfScope prof__(ctx, "softmax_heat");
  dim3 grid(cdiv(wc, 128), hc, n);
  softmax_heat_kernel<<<grid, 128, 0, ctx.stream>>>(reinterpret_cast<const float4*>(semi_c4), c4_total, heat, hc, wc);
  B200M_LAUNCH_CHECK(ctx, "softmax_heat");
}

// ------------------------------------------------------------------------------------------------
// simple_nms (:7-22) + threshold (:135-138) + remove_borders (:25-30), tile-local and exact.
// Five chained (2r+1)^2 max-pools give a 5r-pixel dependency halo; a block evaluates a 32x32 output
// tile on a (32+10r)^2 working window held in shared memory (r <= 4 -> 72x72).  Pools are separable
// (row pass then column pass); positions outside the image act as -inf padding (scores) / 0 (masks),
// positions outside the window are never consumed by a valid output (shrinking-validity argument:
// each pool consumes r pixels of margin, 5 pools consume the 5r halo).
constexpr int kNmsTile = 32;

template <typename T>
__device__ __forceinline__ void pool_pass(const T* __restrict__ src, T* __restrict__ tmp, T* __restrict__ dst,
                                          int Wd, int r, T lowest) {
  const int N = Wd * Wd;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    int y = i / Wd, x = i - y * Wd;
    int lo = max(x - r, 0), hi = min(x + r, Wd - 1);
    T m = lowest;
    for (int k = lo; k <= hi; ++k) { T c = src[y * Wd + k]; m = c > m ? c : m; }
    tmp[i] = m;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    int y = i / Wd, x = i - y * Wd;
    int lo = max(y - r, 0), hi = min(y + r, Wd - 1);
    T m = lowest;
    for (int k = lo; k <= hi; ++k) { T c = tmp[k * Wd + x]; m = c > m ? c : m; }
    dst[i] = m;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) nms_kernel(const float* __restrict__ heat, float* __restrict__ nms_dense,
                                                  int H8, int W8, int r, float thr, int border,
                                                  unsigned long long* __restrict__ cand_keys,
                                                  int* __restrict__ cand_counts, int cand_cap,
                                                  int* __restrict__ overflow_flag) {
  extern __shared__ unsigned char smraw[];
  const int halo = 5 * r;
  const int Wd = kNmsTile + 2 * halo;
  const int N = Wd * Wd;
  float* S = reinterpret_cast<float*>(smraw);   // scores (-inf outside the image)
  float* A = S + N;                             // pooled / suppressed scores
  float* T = A + N;                             // row-pass scratch
  unsigned char* M = reinterpret_cast<unsigned char*>(T + N);   // max_mask
  unsigned char* SP = M + N;                                    // supp_mask
  unsigned char* TB = SP + N;                                   // byte scratch
  unsigned char* IN = TB + N;                                   // inside-image flag
  const int n = blockIdx.z;
  const int x0 = blockIdx.x * kNmsTile - halo, y0 = blockIdx.y * kNmsTile - halo;
  const float* hm = heat + (size_t)n * H8 * W8;
  const float NEG = -INFINITY;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    int y = i / Wd, x = i - y * Wd;
    int gy = y0 + y, gx = x0 + x;
    bool in = gy >= 0 && gy < H8 && gx >= 0 && gx < W8;
    S[i] = in ? hm[(size_t)gy * W8 + gx] : NEG;
    IN[i] = in;
  }
  __syncthreads();
  // max_mask = scores == max_pool(scores)
  pool_pass<float>(S, T, A, Wd, r, NEG);
  for (int i = threadIdx.x; i < N; i += blockDim.x) M[i] = IN[i] && (S[i] == A[i]);
  __syncthreads();
  for (int it = 0; it < 2; ++it) {
    // supp_mask = max_pool(max_mask) > 0
    pool_pass<unsigned char>(M, TB, SP, Wd, r, (unsigned char)0);
    // supp_scores = where(supp_mask, 0, scores)   (padding of the next pool stays -inf)
    for (int i = threadIdx.x; i < N; i += blockDim.x) A[i] = IN[i] ? (SP[i] ? 0.f : S[i]) : NEG;
    __syncthreads();
    // new_max_mask = supp_scores == max_pool(supp_scores): row pass A -> T, column pass fused with the
    // mask update (reads T and the thread's own A[i] only)
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      int y = i / Wd, x = i - y * Wd;
      int lo = max(x - r, 0), hi = min(x + r, Wd - 1);
      float m = NEG;
      for (int k = lo; k <= hi; ++k) m = fmaxf(m, A[y * Wd + k]);
      T[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      int y = i / Wd, x = i - y * Wd;
      int lo = max(y - r, 0), hi = min(y + r, Wd - 1);
      float m = NEG;
      for (int k = lo; k <= hi; ++k) m = fmaxf(m, T[k * Wd + x]);
      // max_mask |= new_max_mask & ~supp_mask
      if (IN[i] && !SP[i] && A[i] == m) M[i] = 1;
    }
    __syncthreads();
  }
  // where(max_mask, scores, 0) -> threshold -> border -> candidate list
  for (int i = threadIdx.x; i < kNmsTile * kNmsTile; i += blockDim.x) {
    int ty = i / kNmsTile, tx = i - ty * kNmsTile;
    int gy = blockIdx.y * kNmsTile + ty, gx = blockIdx.x * kNmsTile + tx;
    if (gy >= H8 || gx >= W8) continue;
    int w = (ty + halo) * Wd + tx + halo;
    float sc = M[w] ? S[w] : 0.f;
    if (nms_dense) nms_dense[(size_t)n * H8 * W8 + (size_t)gy * W8 + gx] = sc;
    if (cand_keys && sc > thr && gy >= border && gy < H8 - border && gx >= border && gx < W8 - border) {
      int slot = atomicAdd(&cand_counts[n], 1);
      if (slot < cand_cap) {
        unsigned int lin = (unsigned int)(gy * W8 + gx);
        // descending sort key: larger score first, then smaller linear index first
        cand_keys[(size_t)n * cand_cap + slot] =
            ((unsigned long long)__float_as_uint(sc) << 32) | (unsigned long long)(0xFFFFFFFFu - lin);
      } else {
        *overflow_flag = 1;
      }
    }
  }
}

static size_t nms_smem_bytes(int r) {
  int Wd = kNmsTile + 10 * r;
  return (size_t)Wd * Wd * (3 * sizeof(float) + 4);
}

void launch_nms_candidates(LaunchCtx& ctx, const float* heat, float* nms_dense, int n, int H8, int W8,
                           int radius, float thr, int border, unsigned long long* cand_keys,
                           int* cand_counts, int cand_cap, int* overflow_flag) {
  ProfScope prof__(ctx, "nms_candidates");
  static size_t attr_bytes = 0;
  size_t bytes = nms_smem_bytes(radius);
  if (bytes > attr_bytes) {
    cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    attr_bytes = bytes;
  }
  dim3 grid(cdiv(W8, kNmsTile), cdiv(H8, kNmsTile), n);
  nms_kernel<<<grid, 256, bytes, ctx.stream>>>(heat, nms_dense, H8, W8, radius, thr, border, cand_keys,
                                              cand_counts, cand_cap, overflow_flag);
  B200M_LAUNCH_CHECK(ctx, "nms");
}

// ------------------------------------------------------------------------------------------------
// top_k_keypoints (:33-37) / row-major order (:135-138): one block per image sorts the candidate
// keys with a bitonic network (shared memory when they fit, the global list otherwise).
//   count >  k >= 0 : descending score (ties: lower linear index first), first k kept
//   otherwise       : ascending linear index (the order torch.nonzero produces)
__device__ void block_bitonic_desc(unsigned long long* keys, int npow2) {
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long a = keys[i], b = keys[ixj];
          bool desc = (i & k) == 0;
          if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
}

constexpr int kSelectSmemKeys = 8192;

__global__ void __launch_bounds__(1024) select_keypoints_kernel(unsigned long long* __restrict__ cand_keys,
                                                                const int* __restrict__ cand_counts,
                                                                int cand_cap, int W8, int max_kp,
                                                                float* __restrict__ keypoints,
                                                                float* __restrict__ scores,
                                                                int* __restrict__ counts, int cap) {
  extern __shared__ unsigned long long skeys[];
  const int n = blockIdx.x;
  unsigned long long* gk = cand_keys + (size_t)n * cand_cap;
  const int cnt = min(cand_counts[n], cand_cap);
  const bool topk = (max_kp >= 0) && (cnt > max_kp);
  int npow2 = 1;
  while (npow2 < cnt) npow2 <<= 1;
  const bool in_smem = npow2 <= kSelectSmemKeys;
  unsigned long long* keys = in_smem ? skeys : gk;   // cand_cap is a power of two >= npow2
  // Build sort keys.  Row-major mode sorts by ~linear-index descending == index ascending.
  for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
    unsigned long long k = 0ull;
    if (i < cnt) {
      k = gk[i];
      if (!topk) k = ((k & 0xFFFFFFFFull) << 32) | (k >> 32);
    }
    keys[i] = k;   // padding keys are 0 -> sort to the end
  }
  __syncthreads();
  block_bitonic_desc(keys, npow2);
  const int keep = topk ? max_kp : min(cnt, cap);
  float* kp = keypoints + (size_t)n * cap * 2;
  float* sc = scores + (size_t)n * cap;
  for (int i = threadIdx.x; i < cap; i += blockDim.x) {
    float x = 0.f, y = 0.f, s = 0.f;
    if (i < keep) {
      unsigned long long k = keys[i];
      if (!topk) k = ((k & 0xFFFFFFFFull) << 32) | (k >> 32);
      unsigned int lin = 0xFFFFFFFFu - (unsigned int)(k & 0xFFFFFFFFull);
      s = __uint_as_float((unsigned int)(k >> 32));
      y = (float)(lin / (unsigned)W8);
      x = (float)(lin % (unsigned)W8);
    }
    kp[2 * i] = x;       // torch.flip(k, [1]).float(): (x, y)   (:151)
    kp[2 * i + 1] = y;
    sc[i] = s;
  }
  if (threadIdx.x == 0) counts[n] = keep;
}

void launch_select_keypoints(LaunchCtx& ctx, unsigned long long* cand_keys, const int* cand_counts,
                             int cand_cap, int n, int W8, int max_kp, float* keypoints, float* scores,
                             int* counts, int cap) {
  ProfScope prof__(ctx, "select_keypoints");
  static bool attr_set = false;
  size_t bytes = (size_t)kSelectSmemKeys * sizeof(unsigned long long);
  if (!attr_set) {
    cudaFuncSetAttribute(select_keypoints_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    attr_set = true;
  }
  select_keypoints_kernel<<<n, 1024, bytes, ctx.stream>>>(cand_keys, cand_counts, cand_cap, W8, max_kp,
                                                        keypoints, scores, counts, cap);
  B200M_LAUNCH_CHECK(ctx, "select_keypoints");
}

// ------------------------------------------------------------------------------------------------
// sample_descriptors (:40-52): bilinear grid_sample (zeros padding) of the normalised descriptor map at
// the keypoints, then L2 normalise (eps 1e-12).  One warp per keypoint, lane = channel group (float4).
__global__ void __launch_bounds__(256) sample_desc_kernel(const float4* __restrict__ desc, int c4_total, int D,
                                                          int hc, int wc, const float* __restrict__ keypoints,
                                                          const int* __restrict__ counts, int cap, int align_corners,
                                                          float* __restrict__ out_dcn, float* __restrict__ out_tok,
                                                          int tok_ld, size_t tok_img_stride) {
  const int n = blockIdx.y;
  const int k = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= cap) return;
  const int cnt = counts ? counts[n] : cap;
  const int G = D / 4;
  if (k >= cnt) {   // zero the padding entries so stacked tensors are deterministic
    for (int g = lane; g < G; g += 32) {
      if (out_dcn)
        for (int j = 0; j < 4; ++j) out_dcn[((size_t)n * D + g * 4 + j) * cap + k] = 0.f;
      if (out_tok)
        reinterpret_cast<float4*>(out_tok + n * tok_img_stride + (size_t)k * tok_ld)[g] = make_float4(0, 0, 0, 0);
    }
    return;
  }
  const float s = 8.f;
  float kx = keypoints[((size_t)n * cap + k) * 2], ky = keypoints[((size_t)n * cap + k) * 2 + 1];
  // keypoints = keypoints - s/2 + 0.5 ; /= [w*s - s/2 - 0.5, h*s - s/2 - 0.5] ; *2 - 1
  float gx = (kx - s / 2 + 0.5f) / ((float)wc * s - s / 2 - 0.5f) * 2.f - 1.f;
  float gy = (ky - s / 2 + 0.5f) / ((float)hc * s - s / 2 - 0.5f) * 2.f - 1.f;
  float px, py;   // grid_sampler unnormalize
  if (align_corners) {
    px = ((gx + 1.f) / 2.f) * (float)(wc - 1);
    py = ((gy + 1.f) / 2.f) * (float)(hc - 1);
  } else {
    px = ((gx + 1.f) * (float)wc - 1.f) / 2.f;
    py = ((gy + 1.f) * (float)hc - 1.f) / 2.f;
  }
  const float fx0 = floorf(px), fy0 = floorf(py);
  const int ix = (int)fx0, iy = (int)fy0;
  // torch's grid_sampler weights: nw = (ix_se - ix)*(iy_se - iy), ne = (ix - ix_sw)*(iy_sw - iy), ...
  const float tx = px - fx0, ty = py - fy0;
  const float ex = (fx0 + 1.f) - px, ey = (fy0 + 1.f) - py;
  const float w00 = ex * ey, w01 = tx * ey, w10 = ex * ty, w11 = tx * ty;
  const size_t plane = (size_t)hc * wc;
  const float4* base = desc + (size_t)n * c4_total * plane;
  float4 acc[2];   // up to D = 256 (64 groups / 32 lanes)
  float ss = 0.f;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    acc[t] = make_float4(0, 0, 0, 0);
    int g = lane + 32 * t;
    if (g < G) {
      const float4* pl = base + (size_t)g * plane;
      auto tap = [&](int yy, int xx, float w) {
        if (yy >= 0 && yy < hc && xx >= 0 && xx < wc) {
          float4 v = pl[(size_t)yy * wc + xx];
          acc[t].x += v.x * w; acc[t].y += v.y * w; acc[t].z += v.z * w; acc[t].w += v.w * w;
        }
      };
      tap(iy, ix, w00);
      tap(iy, ix + 1, w01);
      tap(iy + 1, ix, w10);
      tap(iy + 1, ix + 1, w11);
      ss += acc[t].x * acc[t].x + acc[t].y * acc[t].y + acc[t].z * acc[t].z + acc[t].w * acc[t].w;
    }
  }
  ss = warp_sum(ss);
  const float nrm = fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    int g = lane + 32 * t;
    if (g < G) {
      float4 o = make_float4(acc[t].x / nrm, acc[t].y / nrm, acc[t].z / nrm, acc[t].w / nrm);
      if (out_dcn) {
        float* d = out_dcn + ((size_t)n * D + g * 4) * cap + k;
        d[0] = o.x; d[(size_t)cap] = o.y; d[(size_t)2 * cap] = o.z; d[(size_t)3 * cap] = o.w;
      }
      if (out_tok) reinterpret_cast<float4*>(out_tok + n * tok_img_stride + (size_t)k * tok_ld)[g] = o;
    }
  }
}

void launch_sample_descriptors(LaunchCtx& ctx, const float* desc_c4, int c4_total, int D, int n, int hc, int wc,
                               const float* keypoints, const int* counts, int cap, int align_corners,
                               float* out_dcn, float* out_tok, int tok_ld, size_t tok_img_stride) {
  ProfScope prof__(ctx, "sample_descriptors");
  if (cap <= 0) return;
  dim3 grid(cdiv(cap, 8), n);
  sample_desc_kernel<<<grid, 256, 0, ctx.stream>>>(reinterpret_cast<const float4*>(desc_c4), c4_total, D, hc, wc,
                                                   keypoints, counts, cap, align_corners, out_dcn, out_tok,
                                                   tok_ld, tok_img_stride);
  B200M_LAUNCH_CHECK(ctx, "sample_descriptors");
}

}  // namespace b200m
