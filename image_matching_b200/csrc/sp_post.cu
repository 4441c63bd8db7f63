// SuperPoint detector post-processing and descriptor sampling (HBM / shared-memory bound, fp32,
// compare-only NMS => bit-exact given the same heat-map).
// Reference: superpoint/models/superpoint_test.py:7-52 and :128-155.
#include <stdlib.h>
#include <string.h>
#include "kernels.cuh"
#include "tc_common.cuh"

namespace b200m {

// ------------------------------------------------------------------------------------------------
// softmax over the 65 detector channels, drop the dustbin, depth-to-space x8  (:128-131)
// One thread per coarse cell; a warp covers 32 consecutive cells of a row, so each of the 8 output
// rows receives one 1024 B contiguous run per warp.
__global__ void __launch_bounds__(128) softmax_heat_kernel(const float4* __restrict__ semi, int c4_total,
                                                           float* __restrict__ heat, int hc, int wc) {
  pdl_trigger();   // PDL (common.cuh): the next kernel may be scheduled; nothing is read or written before the wait
  pdl_wait();
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
  launch_pdl(ctx, kPdlPost, softmax_heat_kernel, dim3(grid), dim3(128), 0, reinterpret_cast<const float4*>(semi_c4), c4_total, heat, hc, wc);
  B200M_LAUNCH_CHECK(ctx, "softmax_heat");
}

// ------------------------------------------------------------------------------------------------
// simple_nms (:7-22) + threshold (:135-138) + remove_borders (:25-30), exact (compare-only).
// The reference chains five (2r+1)^2 max-pools.  Evaluating all five inside one tile needs a 5r-pixel halo
// (25x redundant work at 32x32 tiles, measured 12.8 ms / 128 images); instead each pool is its own pass over
// the L2-resident maps with an r-pixel halo (1.56x), separable row / column max in shared memory:
//   pass A  (MODE 0): m      = (s == P(s))
//   pass B  (MODE 1): supp   = P(m) > 0 ;  ss = supp ? 0 : s
//   pass C  (MODE 2): m     |= (ss == P(ss)) & ~supp          (B, C run twice)
//   the last pass C also emits where(m, s, 0), the threshold / border test and the candidate list.
// Out-of-image positions act as -inf padding exactly like torch's max_pool2d.
constexpr int kNmsTile = 32;

template <int R, int MODE, bool LAST>
__global__ void __launch_bounds__(256) nms_pass_kernel(const float* __restrict__ heat, float* __restrict__ ss,
                                                       unsigned char* __restrict__ m, unsigned char* __restrict__ supp,
                                                       float* __restrict__ nms_dense, int H8, int W8, float thr,
                                                       int border, unsigned long long* __restrict__ cand_keys,
                                                       int* __restrict__ cand_counts, int cand_cap,
                                                       int* __restrict__ overflow_flag) {
  constexpr int Wd = kNmsTile + 2 * R;
  __shared__ float in[Wd][Wd + 1];
  __shared__ float tmp[Wd][kNmsTile + 1];
  const int n = blockIdx.z;
  const size_t img = (size_t)n * H8 * W8;
  const int x0 = blockIdx.x * kNmsTile - R, y0 = blockIdx.y * kNmsTile - R;
  const float NEG = -INFINITY;
  // pooled input: scores (A), mask as 0/1 (B), suppressed scores (C)
  for (int i = threadIdx.x; i < Wd * Wd; i += 256) {
    const int y = i / Wd, x = i - y * Wd;
    const int gy = y0 + y, gx = x0 + x;
    float v = NEG;
    if (gy >= 0 && gy < H8 && gx >= 0 && gx < W8) {
      const size_t o = img + (size_t)gy * W8 + gx;
      v = MODE == 0 ? heat[o] : (MODE == 1 ? (float)m[o] : ss[o]);
    }
    in[y][x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Wd * kNmsTile; i += 256) {     // row pass
    const int y = i / kNmsTile, x = i - y * kNmsTile;
    float v = in[y][x];
#pragma unroll
    for (int d = 1; d <= 2 * R; ++d) v = fmaxf(v, in[y][x + d]);
    tmp[y][x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kNmsTile * kNmsTile; i += 256) {   // column pass + combine
    const int y = i / kNmsTile, x = i - y * kNmsTile;
    const int gy = blockIdx.y * kNmsTile + y, gx = blockIdx.x * kNmsTile + x;
    if (gy >= H8 || gx >= W8) continue;
    float v = tmp[y][x];
#pragma unroll
    for (int d = 1; d <= 2 * R; ++d) v = fmaxf(v, tmp[y + d][x]);
    const size_t o = img + (size_t)gy * W8 + gx;
    const float c = in[y + R][x + R];          // the pooled map's own value at this pixel
    if (MODE == 0) {
      m[o] = (c == v);
    } else if (MODE == 1) {
      const bool sp = v > 0.f;
      supp[o] = sp;
      ss[o] = sp ? 0.f : heat[o];
    } else {
      bool mk = m[o] != 0;
      if (!mk && !supp[o] && c == v) mk = true;
      if (!LAST) {
        m[o] = mk;
      } else {
        const float sc = mk ? heat[o] : 0.f;
        if (nms_dense) nms_dense[o] = sc;
        if (cand_keys && sc > thr && gy >= border && gy < H8 - border && gx >= border && gx < W8 - border) {
          const int slot = atomicAdd(&cand_counts[n], 1);
          if (slot < cand_cap) {
            const unsigned int lin = (unsigned int)(gy * W8 + gx);
            // descending sort key: larger score first, then smaller linear index first
            cand_keys[(size_t)n * cand_cap + slot] =
                ((unsigned long long)__float_as_uint(sc) << 32) | (unsigned long long)(0xFFFFFFFFu - lin);
          } else {
            *overflow_flag = 1;
          }
        }
      }
    }
  }
}

template <int R>
static void launch_nms_r(LaunchCtx& ctx, const float* heat, float* ss, unsigned char* m, unsigned char* supp,
                         float* nms_dense, int n, int H8, int W8, float thr, int border,
                         unsigned long long* cand_keys, int* cand_counts, int cand_cap, int* overflow_flag) {
  dim3 grid(cdiv(W8, kNmsTile), cdiv(H8, kNmsTile), n);
#define B200M_NMS_PASS(MODE, LAST)                                                                               \
  nms_pass_kernel<R, MODE, LAST><<<grid, 256, 0, ctx.stream>>>(heat, ss, m, supp, nms_dense, H8, W8, thr, border, \
                                                               cand_keys, cand_counts, cand_cap, overflow_flag);  \
  B200M_LAUNCH_CHECK(ctx, "nms_pass")
  B200M_NMS_PASS(0, false);
  B200M_NMS_PASS(1, false);
  B200M_NMS_PASS(2, false);
  B200M_NMS_PASS(1, false);
  B200M_NMS_PASS(2, true);
#undef B200M_NMS_PASS
}

// ---- nms_radius == 4 (the reference default): all five pools in ONE kernel ---------------------------------------
// A 32x32 output tile with the 20-pixel dependency halo (72x72 scores) stays in shared memory through the five chained
// 9x9 max-pools; pool k is evaluated only where its result can still reach the tile (the valid region shrinks by 4
// pixels = one float4 chunk per pool), so the halo costs 2.5x redundant work instead of 5x.  Each pool is separable and
// evaluated with register sliding windows: a row-pass thread loads three float4 chunks and emits four outputs
// (9 three-input max instructions instead of 32 two-input ones), a column-pass thread walks 12 rows for four outputs.
// The intermediate maps (mask, suppression flags) never leave shared memory: one read of the heat-map, one launch
// (the five-pass version moved ~0.5 GB through the L2 per 16 images and spent 140 us on them; this one ~45 us).
constexpr int kNfHalo = 20;

template <int T>
struct NmsFusedSmem {
  static constexpr int W = T + 2 * kNfHalo;   // 72 (T = 32) or 104 (T = 64): a whole number of float4 chunks
  static constexpr int C = W / 4;
  float s[W][W];            // scores, -inf outside the image
  float t[W][W];            // row-pass result of the current pool
  unsigned char m[W][W];    // max mask
  unsigned char sp[W][W];   // suppression flags of the current round
};

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// SRC: 0 = scores, 1 = mask as 0/1, 2 = suppressed scores (supp ? 0 : s); out-of-image pixels are -inf in all three
template <int SRC, typename SM>
__device__ __forceinline__ float4 nms_src_chunk(const SM& S, int y, int c) {
  const float4 sv = *reinterpret_cast<const float4*>(&S.s[y][4 * c]);
  if (SRC == 0) return sv;
  const float NEG = -INFINITY;
  if (SRC == 1) {
    const uchar4 mv = *reinterpret_cast<const uchar4*>(&S.m[y][4 * c]);
    return make_float4(sv.x == NEG ? NEG : (float)mv.x, sv.y == NEG ? NEG : (float)mv.y,
                       sv.z == NEG ? NEG : (float)mv.z, sv.w == NEG ? NEG : (float)mv.w);
  }
  const uchar4 pv = *reinterpret_cast<const uchar4*>(&S.sp[y][4 * c]);
  return make_float4(pv.x && sv.x != NEG ? 0.f : sv.x, pv.y && sv.y != NEG ? 0.f : sv.y,
                     pv.z && sv.z != NEG ? 0.f : sv.z, pv.w && sv.w != NEG ? 0.f : sv.w);
}

// pool number K (0..4) of the chain: input valid on [4K, W-4K)^2, output on [4K+4, W-4K-4)^2
template <int K, int SRC, int NT, typename SM>
__device__ __forceinline__ void nms_row_pass(SM& S) {
  constexpr int rows = SM::W - 8 * K, c0 = K + 1, nc = SM::C - 2 * (K + 1);
#pragma unroll 2
  for (int i = threadIdx.x; i < rows * nc; i += NT) {
    const int y = 4 * K + i / nc, c = c0 + i % nc;
    const float4 a = nms_src_chunk<SRC>(S, y, c - 1), b = nms_src_chunk<SRC>(S, y, c), d = nms_src_chunk<SRC>(S, y, c + 1);
    // inputs a.x..a.w b.x..b.w d.x..d.w = positions -4..7 relative to the chunk; output j covers positions j-4..j+4
    const float core = fmaxf(max3(a.w, b.x, b.y), max3(b.z, b.w, d.x));     // positions -1..4, common to all four
    float4 o;
    o.x = fmaxf(max3(a.x, a.y, a.z), core);                                  // -4..4
    o.y = fmaxf(max3(a.y, a.z, d.y), core);                                  // -3..5
    o.z = fmaxf(max3(a.z, d.y, d.z), core);                                  // -2..6
    o.w = fmaxf(max3(d.y, d.z, d.w), core);                                  // -1..7
    *reinterpret_cast<float4*>(&S.t[y][4 * c]) = o;
  }
}

// column pass + the element-wise step that follows pool K.  EP: 0 -> m = (s == P(s));  1 -> supp = P(m) > 0;
// 2 -> m |= (ss == P(ss)) & ~supp.  `fin` receives (y, x, keep, score) for the last pool.
template <int K, int EP, int NT, typename SM, typename Fin>
__device__ __forceinline__ void nms_col_pass(SM& S, Fin fin) {
  constexpr int lo = 4 * (K + 1), w = SM::W - 8 * (K + 1), g0 = K + 1, ng = SM::C - 2 * (K + 1);
  const float NEG = -INFINITY;
#pragma unroll 2
  for (int i = threadIdx.x; i < w * ng; i += NT) {
    const int x = lo + i % w, g = g0 + i / w;
    float v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = S.t[4 * g - 4 + k][x];
    const float core = fmaxf(max3(v[3], v[4], v[5]), max3(v[6], v[7], v[8]));
    float o[4];
    o[0] = fmaxf(max3(v[0], v[1], v[2]), core);
    o[1] = fmaxf(max3(v[1], v[2], v[9]), core);
    o[2] = fmaxf(max3(v[2], v[9], v[10]), core);
    o[3] = fmaxf(max3(v[9], v[10], v[11]), core);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int y = 4 * g + j;
      const float sc = S.s[y][x];
      const bool inside = sc != NEG;
      if (EP == 0) {
        S.m[y][x] = inside && sc == o[j];
      } else if (EP == 1) {
        S.sp[y][x] = o[j] > 0.f;
      } else {
        const bool sup = S.sp[y][x] != 0;
        const float ss = sup ? 0.f : sc;
        const bool mk = S.m[y][x] != 0 || (inside && !sup && ss == o[j]);
        if (K < 4) S.m[y][x] = mk;
        else fin(y, x, mk, sc);
      }
    }
  }
}

template <int T, int NT>
__global__ void __launch_bounds__(NT) nms_fused_r4_kernel(const float* __restrict__ heat, float* __restrict__ nms_dense,
                                                          int H8, int W8, float thr, int border,
                                                          unsigned long long* __restrict__ cand_keys,
                                                          int* __restrict__ cand_counts, int cand_cap,
                                                          int* __restrict__ overflow_flag) {
  pdl_trigger();   // PDL (common.cuh): the next kernel may be scheduled; nothing is read or written before the wait
  pdl_wait();
  using SM = NmsFusedSmem<T>;
  extern __shared__ __align__(16) unsigned char nms_smem_raw[];
  SM& S = *reinterpret_cast<SM*>(nms_smem_raw);
  const int n = blockIdx.z;
  const size_t img = (size_t)n * H8 * W8;
  const int x0 = blockIdx.x * T - kNfHalo, y0 = blockIdx.y * T - kNfHalo;   // multiples of 4; W8 is one too
  const float NEG = -INFINITY;
  {
    // the whole halo tile in one go: all of a thread's loads are issued before the first shared-memory store
    constexpr int kItems = SM::W * SM::C, kPer = (kItems + NT - 1) / NT;
    float4 v[kPer];
#pragma unroll
    for (int r = 0; r < kPer; ++r) {
      const int i = threadIdx.x + r * NT;
      const int y = i / SM::C, c = i - y * SM::C;
      const int gy = y0 + y, gx = x0 + 4 * c;
      v[r] = make_float4(NEG, NEG, NEG, NEG);
      if (i < kItems && gy >= 0 && gy < H8 && gx >= 0 && gx < W8)
        v[r] = __ldg(reinterpret_cast<const float4*>(heat + img + (size_t)gy * W8 + gx));
    }
#pragma unroll
    for (int r = 0; r < kPer; ++r) {
      const int i = threadIdx.x + r * NT;
      const int y = i / SM::C, c = i - y * SM::C;
      if (i < kItems) *reinterpret_cast<float4*>(&S.s[y][4 * c]) = v[r];
    }
  }
  auto none = [](int, int, bool, float) {};
  __syncthreads();
  nms_row_pass<0, 0, NT>(S); __syncthreads(); nms_col_pass<0, 0, NT>(S, none); __syncthreads();   // m = s == P(s)
  nms_row_pass<1, 1, NT>(S); __syncthreads(); nms_col_pass<1, 1, NT>(S, none); __syncthreads();   // supp = P(m) > 0
  nms_row_pass<2, 2, NT>(S); __syncthreads(); nms_col_pass<2, 2, NT>(S, none); __syncthreads();   // m |= ...
  nms_row_pass<3, 1, NT>(S); __syncthreads(); nms_col_pass<3, 1, NT>(S, none); __syncthreads();
  nms_row_pass<4, 2, NT>(S); __syncthreads();
  nms_col_pass<4, 2, NT>(S, [&](int y, int x, bool mk, float s) {
    const int gy = y0 + y, gx = x0 + x;
    if (gy >= H8 || gx >= W8) return;
    const float sc = mk ? s : 0.f;
    const size_t o = img + (size_t)gy * W8 + gx;
    if (nms_dense) nms_dense[o] = sc;
    if (cand_keys && sc > thr && gy >= border && gy < H8 - border && gx >= border && gx < W8 - border) {
      const int slot = atomicAdd(&cand_counts[n], 1);
      if (slot < cand_cap) {
        const unsigned int lin = (unsigned int)(gy * W8 + gx);
        cand_keys[(size_t)n * cand_cap + slot] =
            ((unsigned long long)__float_as_uint(sc) << 32) | (unsigned long long)(0xFFFFFFFFu - lin);
      } else {
        *overflow_flag = 1;
      }
    }
  });
}

template <int T, int NT>
static void launch_nms_fused_r4_t(LaunchCtx& ctx, const float* heat, float* nms_dense, int n, int H8, int W8, float thr,
                                  int border, unsigned long long* cand_keys, int* cand_counts, int cand_cap,
                                  int* overflow_flag) {
  static SmemOptIn opt;
  auto kern = nms_fused_r4_kernel<T, NT>;
  opt.ensure(kern, (int)sizeof(NmsFusedSmem<T>));
  dim3 grid(cdiv(W8, T), cdiv(H8, T), n);
  launch_pdl(ctx, kPdlPost, kern, grid, dim3(NT), sizeof(NmsFusedSmem<T>), heat, nms_dense, H8, W8, thr, border, cand_keys,
             cand_counts, cand_cap, overflow_flag);
  B200M_LAUNCH_CHECK(ctx, "nms_fused");
}

// 64x64 tiles (2.6x halo redundancy, 108 KB, two 512-thread CTAs per SM) for real images; 32x32 tiles (5x, 52 KB) when the
// image is so small that 64x64 tiles would leave most of the chip idle
static void launch_nms_fused_r4(LaunchCtx& ctx, const float* heat, float* nms_dense, int n, int H8, int W8, float thr,
                                int border, unsigned long long* cand_keys, int* cand_counts, int cand_cap,
                                int* overflow_flag) {
  static const int forced = [] { const char* e = getenv("B200M_NMS_TILE"); return e ? atoi(e) : 0; }();
  const bool big = forced ? forced == 64 : (size_t)cdiv(W8, 64) * cdiv(H8, 64) * n >= 296;
  if (big)
    launch_nms_fused_r4_t<64, 512>(ctx, heat, nms_dense, n, H8, W8, thr, border, cand_keys, cand_counts, cand_cap,
                                   overflow_flag);
  else
    launch_nms_fused_r4_t<32, 256>(ctx, heat, nms_dense, n, H8, W8, thr, border, cand_keys, cand_counts, cand_cap,
                                   overflow_flag);
}

size_t nms_scratch_bytes(int n, int H8, int W8) { return (size_t)n * H8 * W8 * 6 + 1024; }

void launch_nms_candidates(LaunchCtx& ctx, const float* heat, float* nms_dense, int n, int H8, int W8,
                           int radius, float thr, int border, unsigned long long* cand_keys,
                           int* cand_counts, int cand_cap, int* overflow_flag, void* scratch) {
  ProfScope prof__(ctx, "nms_candidates");
  static const bool multipass = [] { const char* e = getenv("B200M_NMS_IMPL"); return e && strcmp(e, "multipass") == 0; }();
  if (radius == 4 && W8 % 4 == 0 && !multipass) {
    launch_nms_fused_r4(ctx, heat, nms_dense, n, H8, W8, thr, border, cand_keys, cand_counts, cand_cap, overflow_flag);
    return;
  }
  const size_t px = (size_t)n * H8 * W8;
  float* ss = reinterpret_cast<float*>(scratch);
  unsigned char* m = reinterpret_cast<unsigned char*>(ss + px);
  unsigned char* supp = m + px;
  switch (radius) {
    case 0: launch_nms_r<0>(ctx, heat, ss, m, supp, nms_dense, n, H8, W8, thr, border, cand_keys, cand_counts, cand_cap, overflow_flag); break;
    case 1: launch_nms_r<1>(ctx, heat, ss, m, supp, nms_dense, n, H8, W8, thr, border, cand_keys, cand_counts, cand_cap, overflow_flag); break;
    case 2: launch_nms_r<2>(ctx, heat, ss, m, supp, nms_dense, n, H8, W8, thr, border, cand_keys, cand_counts, cand_cap, overflow_flag); break;
    case 3: launch_nms_r<3>(ctx, heat, ss, m, supp, nms_dense, n, H8, W8, thr, border, cand_keys, cand_counts, cand_cap, overflow_flag); break;
    default: launch_nms_r<4>(ctx, heat, ss, m, supp, nms_dense, n, H8, W8, thr, border, cand_keys, cand_counts, cand_cap, overflow_flag); break;
  }
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

constexpr int kSelectSmemKeys = 16384;   // 128 KB: the ~10 k candidates of a 1280x960 image sort in shared memory

__global__ void __launch_bounds__(1024) select_keypoints_kernel(unsigned long long* __restrict__ cand_keys,
                                                                const int* __restrict__ cand_counts,
                                                                int cand_cap, int W8, int max_kp,
                                                                float* __restrict__ keypoints,
                                                                float* __restrict__ scores,
                                                                int* __restrict__ counts, int cap) {
  pdl_trigger();   // PDL (common.cuh): the next kernel may be scheduled; nothing is read or written before the wait
  pdl_wait();
  extern __shared__ unsigned long long skeys[];
  const int n = blockIdx.x;
  unsigned long long* gk = cand_keys + (size_t)n * cand_cap;
  const int cnt = min(cand_counts[n], cand_cap);
  const bool topk = (max_kp >= 0) && (cnt > max_kp);
  int npow2 = 1;
  while (npow2 < cnt) npow2 <<= 1;
  const bool in_smem = npow2 <= kSelectSmemKeys;
  unsigned long long* keys = in_smem ? skeys : gk;   // cand_cap is a power of two >= npow2
  int nsort = npow2;
  if (topk && !in_smem && max_kp <= kSelectSmemKeys) {
    // More candidates than shared memory holds (a 1280x960 image easily has 30 k): instead of a bitonic sort of all of
    // them in global memory, find the max_kp-th largest key with an MSB radix select (8 passes over the L2-resident
    // list, 256-bin shared-memory histograms; the keys are unique, so exactly max_kp keys are >= it), compact those
    // into shared memory and sort only them.
    __shared__ int hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_remaining, s_fill;
    if (threadIdx.x == 0) { s_prefix = 0ull; s_remaining = max_kp; s_fill = 0; }
    for (int shift = 56; shift >= 0; shift -= 8) {
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const unsigned long long k = gk[i];
        const bool match = shift == 56 || (k >> (shift + 8)) == (prefix >> (shift + 8));
        if (match) atomicAdd(&hist[(int)((k >> shift) & 255ull)], 1);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int acc = 0, d = 255;
        for (; d > 0; --d) {
          if (acc + hist[d] >= s_remaining) break;
          acc += hist[d];
        }
        s_remaining -= acc;                       // rank of the wanted key inside bin d
        s_prefix = prefix | ((unsigned long long)d << shift);
      }
      __syncthreads();
    }
    const unsigned long long kth = s_prefix;
    nsort = 1;
    while (nsort < max_kp) nsort <<= 1;
    for (int i = threadIdx.x; i < nsort; i += blockDim.x) skeys[i] = 0ull;
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const unsigned long long k = gk[i];
      if (k >= kth) skeys[atomicAdd(&s_fill, 1)] = k;
    }
    keys = skeys;
  } else {
    // Build sort keys.  Row-major mode sorts by ~linear-index descending == index ascending.
    for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
      unsigned long long k = 0ull;
      if (i < cnt) {
        k = gk[i];
        if (!topk) k = ((k & 0xFFFFFFFFull) << 32) | (k >> 32);
      }
      keys[i] = k;   // padding keys are 0 -> sort to the end
    }
  }
  __syncthreads();
  block_bitonic_desc(keys, nsort);
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
  static SmemOptIn opt;
  size_t bytes = (size_t)kSelectSmemKeys * sizeof(unsigned long long);
  opt.ensure(select_keypoints_kernel, (int)bytes);
  launch_pdl(ctx, kPdlPost, select_keypoints_kernel, dim3(n), dim3(1024), bytes, cand_keys, cand_counts, cand_cap, W8, max_kp,
                                                        keypoints, scores, counts, cap);
  B200M_LAUNCH_CHECK(ctx, "select_keypoints");
}

// Sticky error flags (candidate-list overflow, fp16 activation overflow) are surfaced through the per-image counts
// the caller reads back anyway: counts[i] = -1 / -2, so no extra host synchronisation is needed.
__global__ void apply_flags_kernel(const int* __restrict__ flags, int* __restrict__ counts, int n) {
  pdl_trigger();   // PDL (common.cuh): the next kernel may be scheduled; nothing is read or written before the wait
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flags[1]) counts[i] = -2;
  else if (flags[0]) counts[i] = -1;
}

void launch_apply_flags(LaunchCtx& ctx, const int* flags, int* counts, int n) {
  ProfScope prof__(ctx, "apply_flags");
  launch_pdl(ctx, kPdlPost, apply_flags_kernel, dim3(cdiv(n, 128)), dim3(128), 0, flags, counts, n);
  B200M_LAUNCH_CHECK(ctx, "apply_flags");
}

// ------------------------------------------------------------------------------------------------
// sample_descriptors (:40-52): bilinear grid_sample (zeros padding) of the normalised descriptor map at
// the keypoints, then L2 normalise (eps 1e-12).  One warp per keypoint, lane = channel group (float4).
__global__ void __launch_bounds__(256) sample_desc_kernel(const float4* __restrict__ desc, int c4_total, int D,
                                                          int hc, int wc, const float* __restrict__ keypoints,
                                                          const int* __restrict__ counts, int cap, int align_corners,
                                                          float* __restrict__ out_dcn, float* __restrict__ out_tok,
                                                          int tok_ld, size_t tok_img_stride,
                                                          const float* __restrict__ sumsq, int ncb) {
  pdl_trigger();   // PDL (common.cuh): the next kernel may be scheduled; nothing is read or written before the wait
  pdl_wait();
  // channel-major output (B, D, cap): a warp owns ONE keypoint, so writing it directly puts 4 bytes into each of D
  // different sectors.  With cap % 8 == 0 the block's 8 keypoints are transposed through shared memory instead and every
  // channel row leaves as one 32-byte store (a full sector): the kernel ran at 28 % of the HBM rate because of the
  // partial-sector writes.
  __shared__ __align__(16) float s_t[8][256];           // [keypoint of this block][channel], D <= 256
  const bool staged = out_dcn != nullptr && (cap & 7) == 0 && D <= 256 &&
                      (reinterpret_cast<uintptr_t>(out_dcn) & 31) == 0;      // uniform per launch
  const int n = blockIdx.y;
  const int wk = threadIdx.x >> 5;
  const int k = blockIdx.x * (blockDim.x / 32) + wk;
  const int lane = threadIdx.x & 31;
  const int G = D / 4;
  auto flush = [&]() {                                   // all 256 threads
    __syncthreads();
    const int k0 = blockIdx.x * 8;
    if ((int)threadIdx.x < D && k0 < cap) {
      uint32_t r[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) r[q] = __float_as_uint(s_t[q][threadIdx.x]);
      tc::st_global_256(out_dcn + ((size_t)n * D + threadIdx.x) * cap + k0, r);
    }
  };
  if (k >= cap) {
    if (staged) flush();
    return;
  }
  const int cnt = counts ? counts[n] : cap;
  if (k >= cnt) {   // zero the padding entries so stacked tensors are deterministic
    for (int g = lane; g < G; g += 32) {
      if (staged) {
        *reinterpret_cast<float4*>(&s_t[wk][g * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else if (out_dcn) {
        for (int j = 0; j < 4; ++j) out_dcn[((size_t)n * D + g * 4 + j) * cap + k] = 0.f;
      }
      if (out_tok)
        reinterpret_cast<float4*>(out_tok + n * tok_img_stride + (size_t)k * tok_ld)[g] = make_float4(0, 0, 0, 0);
    }
    if (staged) flush();
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
  // raw head output: the channel norm of each of the four source pixels (superpoint_test.py:125-126,
  // desc / torch.norm(desc, p=2, dim=1)), from the partial sums of squares the head's epilogue wrote
  float nrm4[4] = {1.f, 1.f, 1.f, 1.f};
  if (sumsq) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int yy = iy + (q >> 1), xx = ix + (q & 1);
      if (yy >= 0 && yy < hc && xx >= 0 && xx < wc) {
        float ssq = 0.f;
        for (int cb = 0; cb < ncb; ++cb) ssq += sumsq[((size_t)(n * ncb + cb) * hc + yy) * wc + xx];
        nrm4[q] = sqrtf(ssq);
      }
    }
  }
  float4 acc[2];   // up to D = 256 (64 groups / 32 lanes)
  float ss = 0.f;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    acc[t] = make_float4(0, 0, 0, 0);
    int g = lane + 32 * t;
    if (g < G) {
      const float4* pl = base + (size_t)g * plane;
      auto tap = [&](int yy, int xx, float w, float nrm) {
        if (yy >= 0 && yy < hc && xx >= 0 && xx < wc) {
          float4 v = pl[(size_t)yy * wc + xx];
          if (sumsq) { v.x = v.x / nrm; v.y = v.y / nrm; v.z = v.z / nrm; v.w = v.w / nrm; }   // IEEE division, as the dense normalisation did
          acc[t].x += v.x * w; acc[t].y += v.y * w; acc[t].z += v.z * w; acc[t].w += v.w * w;
        }
      };
      tap(iy, ix, w00, nrm4[0]);
      tap(iy, ix + 1, w01, nrm4[1]);
      tap(iy + 1, ix, w10, nrm4[2]);
      tap(iy + 1, ix + 1, w11, nrm4[3]);
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
      if (staged) {
        *reinterpret_cast<float4*>(&s_t[wk][g * 4]) = o;
      } else if (out_dcn) {
        float* d = out_dcn + ((size_t)n * D + g * 4) * cap + k;
        d[0] = o.x; d[(size_t)cap] = o.y; d[(size_t)2 * cap] = o.z; d[(size_t)3 * cap] = o.w;
      }
      if (out_tok) reinterpret_cast<float4*>(out_tok + n * tok_img_stride + (size_t)k * tok_ld)[g] = o;
    }
  }
  if (staged) flush();
}

void launch_sample_descriptors(LaunchCtx& ctx, const float* desc_c4, int c4_total, int D, int n, int hc, int wc,
                               const float* keypoints, const int* counts, int cap, int align_corners,
                               float* out_dcn, float* out_tok, int tok_ld, size_t tok_img_stride,
                               const float* sumsq, int ncb) {
  ProfScope prof__(ctx, "sample_descriptors");
  if (cap <= 0) return;
  dim3 grid(cdiv(cap, 8), n);
  launch_pdl(ctx, kPdlPost, sample_desc_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<const float4*>(desc_c4), c4_total, D, hc, wc,
                                                   keypoints, counts, cap, align_corners, out_dcn, out_tok,
                                                   tok_ld, tok_img_stride, sumsq, ncb);
  B200M_LAUNCH_CHECK(ctx, "sample_descriptors");
}

}  // namespace b200m
