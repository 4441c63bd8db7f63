// Registration step that follows Matching.forward in the reference's caller (superpoint_glue_test.py:83-92,101):
//   mkpts0 = kpts0[matches > -1]; mkpts1 = kpts1[matches[valid]]
//   Matrix, mask = cv2.estimateAffinePartial2D(mkpts0, mkpts1, method=cv2.RANSAC, ransacReprojThreshold=7)
//   Transform = cv2.warpAffine(source_original, Matrix, (w, h))
// The arithmetic lives in OpenCV (not vendored by the reference; 4.13.0 in this image).  Its published algorithm is
// restated here so that the inlier mask is IDENTICAL to cv2's and the matrix agrees to ~1e-12:
//   * RANSACPointSetRegistrator::run with modelPoints = 2: RNG(-1) multiply-with-carry stream, subsets drawn with
//     uniform(0, count) and re-drawn on duplicates, the closed-form 2-point similarity (doubles), reprojection errors
//     in fp32 (no FMA contraction), `err <= (float)(thr*thr)`, strict improvement of the inlier count and the adaptive
//     iteration bound RANSACUpdateNumIters(confidence, outlier ratio, 2, niters).
//   * the refinement cv2 runs afterwards (10 Levenberg-Marquardt iterations on the inliers' reprojection error) is a
//     LINEAR least-squares problem for the 4-parameter similarity, so its fixed point is the closed-form solution
//     computed here (measured |dM| <= 2e-13 against cv2 over 300 random problems).
//   * warpAffine: inverse map in doubles, 10-bit fixed-point coordinates with a 1/32-pixel interpolation grid
//     (AB_BITS = 10, INTER_BITS = 5), bilinear weights in fp32 (integer 15-bit weights for 8-bit images), constant 0
//     border -- bit-identical to cv2 for uint8 / float32 / float64 images.
// One CTA per image pair for the estimator: the sequential part of RANSAC (RNG stream, adaptive stop) costs nothing,
// the hypothesis scoring is spread over the CTA's 16 warps, one hypothesis per warp per round.
#include "kernels.cuh"

namespace b200m {

namespace {

constexpr int kRansacThreads = 512;
constexpr int kRansacWarps = kRansacThreads / 32;

struct CvRng {   // cv::RNG (multiply-with-carry), core/operations.hpp
  unsigned long long state;
  __device__ unsigned next() {
    state = (unsigned long long)(unsigned)state * 4164903690ULL + (unsigned)(state >> 32);
    return (unsigned)state;
  }
  __device__ int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

// AffinePartial2DEstimatorCallback::runKernel: the similarity through two correspondences (doubles, unfused)
__device__ void similarity_from_two(float2 f0, float2 f1, float2 t0, float2 t1, double* M) {
  const double x1 = f0.x, y1 = f0.y, x2 = f1.x, y2 = f1.y;
  const double X1 = t0.x, Y1 = t0.y, X2 = t1.x, Y2 = t1.y;
  const double dx = __dsub_rn(x1, x2), dy = __dsub_rn(y1, y2);
  const double dX = __dsub_rn(X1, X2), dY = __dsub_rn(Y1, Y2);
  const double d = __ddiv_rn(1.0, __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
  const double cxy = __dsub_rn(__dmul_rn(x1, y2), __dmul_rn(x2, y1));
  const double S0 = __dmul_rn(d, __dadd_rn(__dmul_rn(dX, dx), __dmul_rn(dY, dy)));
  const double S1 = __dmul_rn(d, __dsub_rn(__dmul_rn(dY, dx), __dmul_rn(dX, dy)));
  const double S2 = __dmul_rn(d, __dsub_rn(__dsub_rn(__dmul_rn(dY, cxy),
                                                     __dmul_rn(__dsub_rn(__dmul_rn(X1, y2), __dmul_rn(X2, y1)), dy)),
                                           __dmul_rn(__dsub_rn(__dmul_rn(X1, x2), __dmul_rn(X2, x1)), dx)));
  const double S3 = __dmul_rn(d, __dsub_rn(__dsub_rn(__dmul_rn(-dX, cxy),
                                                     __dmul_rn(__dsub_rn(__dmul_rn(Y1, x2), __dmul_rn(Y2, x1)), dx)),
                                           __dmul_rn(__dsub_rn(__dmul_rn(Y1, y2), __dmul_rn(Y2, y1)), dy)));
  M[0] = S0; M[1] = -S1; M[2] = S2; M[3] = S1; M[4] = S0; M[5] = S3;
}

// Affine2DEstimatorCallback::computeError for one correspondence (fp32, left to right, unfused)
__device__ __forceinline__ float reproj_err(const float* F, float2 f, float2 t) {
  const float a = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(F[0], f.x), __fmul_rn(F[1], f.y)), F[2]), t.x);
  const float b = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(F[3], f.x), __fmul_rn(F[4], f.y)), F[5]), t.y);
  return __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
}

// cv::RANSACUpdateNumIters (calib3d/ptsetreg.cpp) with modelPoints = 2
__device__ int ransac_update_iters(double p, double ep, int max_iters) {
  p = fmin(fmax(p, 0.), 1.);
  ep = fmin(fmax(ep, 0.), 1.);
  const double kDblMin = 2.2250738585072014e-308;
  double num = fmax(1. - p, kDblMin);
  const double q = 1. - ep;
  double denom = 1. - q * q;
  if (denom < kDblMin) return 0;
  num = log(num);
  denom = log(denom);
  return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : __double2int_rn(num / denom);
}

__device__ double block_sum(double v, double* red) {   // fixed-order tree: deterministic
  const int t = threadIdx.x;
  __syncthreads();
  red[t] = v;
  __syncthreads();
  for (int s = kRansacThreads / 2; s > 0; s >>= 1) {
    if (t < s) red[t] += red[t + s];
    __syncthreads();
  }
  return red[0];
}

// info (B,4) int32: [0] correspondences (matches0 > -1), [1] inliers, [2] RANSAC iterations run, [3] 1 if a model
// was found (cv2 returns None / an all-zero mask otherwise)
__global__ void __launch_bounds__(kRansacThreads)
ransac_affine_partial_kernel(const float2* __restrict__ kpts0, const float2* __restrict__ kpts1,
                             const long long* __restrict__ matches0, const int* __restrict__ counts0, int N, int M,
                             double thr, int max_iters, double confidence, int refine, double* __restrict__ matrices,
                             unsigned char* __restrict__ inlier0, int* __restrict__ info) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* from = reinterpret_cast<float2*>(smem_raw);
  float2* to = from + N;
  int* orig = reinterpret_cast<int*>(to + N);
  __shared__ double red[kRansacThreads];
  __shared__ int scan[kRansacThreads];
  __shared__ int sub[kRansacWarps][2];
  __shared__ double modelW[kRansacWarps][6];
  __shared__ int goodW[kRansacWarps];
  __shared__ double best_model[6];
  __shared__ int s_count, s_done, s_best, s_iter;

  const int b = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int n = counts0 ? min(counts0[b], N) : N;
  kpts0 += (size_t)b * N;
  kpts1 += (size_t)b * M;
  matches0 += (size_t)b * N;
  inlier0 += (size_t)b * N;
  matrices += (size_t)b * 6;
  info += b * 4;

  // ---- mkpts0 = kpts0[valid], mkpts1 = kpts1[matches[valid]] (order preserved) ----------------
  const int per = (N + kRansacThreads - 1) / kRansacThreads;
  const int lo = min(t * per, N), hi = min(lo + per, N);
  int c = 0;
  for (int i = lo; i < hi; ++i) {
    inlier0[i] = 0;
    const long long m = i < n ? matches0[i] : -1;
    c += (m > -1 && m < M);
  }
  scan[t] = c;
  __syncthreads();
  for (int s = 1; s < kRansacThreads; s <<= 1) {   // Hillis-Steele inclusive scan
    const int v = t >= s ? scan[t - s] : 0;
    __syncthreads();
    scan[t] += v;
    __syncthreads();
  }
  int pos = scan[t] - c;
  if (t == kRansacThreads - 1) s_count = scan[t];
  for (int i = lo; i < hi; ++i) {
    const long long m = i < n ? matches0[i] : -1;
    if (m > -1 && m < M) { from[pos] = kpts0[i]; to[pos] = kpts1[m]; orig[pos] = i; ++pos; }
  }
  __syncthreads();
  const int count = s_count;

  if (count < 2) {   // cv2: result = false -> H released, mask zeros
    if (t < 6) matrices[t] = 0.0;
    if (t == 0) { info[0] = count; info[1] = 0; info[2] = 0; info[3] = 0; }
    return;
  }
  if (count == 2) {  // exact model through the two points, mask all ones, no refinement
    if (t == 0) {
      double Mm[6];
      similarity_from_two(from[0], from[1], to[0], to[1], Mm);
      for (int k = 0; k < 6; ++k) matrices[k] = Mm[k];
      inlier0[orig[0]] = 1; inlier0[orig[1]] = 1;
      info[0] = 2; info[1] = 2; info[2] = 0; info[3] = 1;
    }
    return;
  }

  // ---- RANSAC: one hypothesis per warp per round; thread 0 replays cv2's sequential bookkeeping ----
  const float thr2 = (float)(thr * thr);
  CvRng rng; rng.state = 0xFFFFFFFFFFFFFFFFULL;   // RNG rng((uint64)-1)
  int niters = max(max_iters, 1);
  if (t == 0) { s_done = 0; s_best = 0; s_iter = 0; }
  __syncthreads();
  while (true) {
    if (t == 0) {
      for (int w = 0; w < kRansacWarps; ++w) {   // getSubset: re-draw on duplicates; 2 points are never degenerate
        const int i0 = rng.uniform(0, count);
        int i1 = rng.uniform(0, count);
        while (i1 == i0) i1 = rng.uniform(0, count);
        sub[w][0] = i0; sub[w][1] = i1;
      }
    }
    __syncthreads();
    {
      double Mm[6];
      similarity_from_two(from[sub[warp][0]], from[sub[warp][1]], to[sub[warp][0]], to[sub[warp][1]], Mm);
      float F[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) F[k] = __double2float_rn(Mm[k]);
      int good = 0;
      for (int i = lane; i < count; i += 32) good += reproj_err(F, from[i], to[i]) <= thr2;
      good = __reduce_add_sync(0xffffffffu, good);
      if (lane == 0) {
        goodW[warp] = good;
#pragma unroll
        for (int k = 0; k < 6; ++k) modelW[warp][k] = Mm[k];
      }
    }
    __syncthreads();
    if (t == 0) {
      int iter = s_iter, best = s_best;
      for (int w = 0; w < kRansacWarps && iter < niters; ++w, ++iter) {
        if (goodW[w] > max(best, 1)) {
          best = goodW[w];
          for (int k = 0; k < 6; ++k) best_model[k] = modelW[w][k];
          niters = ransac_update_iters(confidence, (double)(count - best) / count, niters);
        }
      }
      s_iter = iter; s_best = best;
      s_done = iter >= niters;
    }
    __syncthreads();
    if (s_done) break;
  }
  const int best = s_best;
  if (best <= 0) {
    if (t < 6) matrices[t] = 0.0;
    if (t == 0) { info[0] = count; info[1] = 0; info[2] = s_iter; info[3] = 0; }
    return;
  }

  // ---- inlier mask of the best model (same arithmetic -> same set), then the least-squares refinement ----
  float F[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) F[k] = __double2float_rn(best_model[k]);
  double sx = 0, sy = 0, sX = 0, sY = 0;
  for (int i = t; i < count; i += kRansacThreads) {
    const bool in = reproj_err(F, from[i], to[i]) <= thr2;
    if (in) {
      inlier0[orig[i]] = 1;
      sx += from[i].x; sy += from[i].y; sX += to[i].x; sY += to[i].y;
    } else {
      from[i].x = nanf("");   // marks outliers for the second pass
    }
  }
  double Mo[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) Mo[k] = best_model[k];
  if (refine) {
    const double inv = 1.0 / best;
    const double mx = block_sum(sx, red) * inv, my = block_sum(sy, red) * inv;
    const double mX = block_sum(sX, red) * inv, mY = block_sum(sY, red) * inv;
    double den = 0, na = 0, nb = 0;
    for (int i = t; i < count; i += kRansacThreads) {
      if (from[i].x == from[i].x) {
        const double xc = from[i].x - mx, yc = from[i].y - my, Xc = to[i].x - mX, Yc = to[i].y - mY;
        den += xc * xc + yc * yc;
        na += xc * Xc + yc * Yc;
        nb += xc * Yc - yc * Xc;
      }
    }
    den = block_sum(den, red); na = block_sum(na, red); nb = block_sum(nb, red);
    if (den > 0) {
      const double a = na / den, bb = nb / den;
      Mo[0] = a; Mo[1] = -bb; Mo[2] = mX - (a * mx - bb * my);
      Mo[3] = bb; Mo[4] = a; Mo[5] = mY - (bb * mx + a * my);
    }
  }
  if (t < 6) matrices[t] = Mo[t];
  if (t == 0) { info[0] = count; info[1] = best; info[2] = s_iter; info[3] = 1; }
}

// ------------------------------------------------------------------------------------------ warpAffine
template <typename T> struct WarpAcc;
template <> struct WarpAcc<float> {
  static __device__ __forceinline__ float run(float v0, float v1, float v2, float v3, int ax, int ay) {
    const float fx = __fmul_rn((float)ax, 0.03125f), fy = __fmul_rn((float)ay, 0.03125f);
    const float gx = __fsub_rn(1.f, fx), gy = __fsub_rn(1.f, fy);
    const float w0 = __fmul_rn(gy, gx), w1 = __fmul_rn(gy, fx), w2 = __fmul_rn(fy, gx), w3 = __fmul_rn(fy, fx);
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v0, w0), __fmul_rn(v1, w1)), __fmul_rn(v2, w2)), __fmul_rn(v3, w3));
  }
};
template <> struct WarpAcc<double> {
  static __device__ __forceinline__ double run(double v0, double v1, double v2, double v3, int ax, int ay) {
    const float fx = __fmul_rn((float)ax, 0.03125f), fy = __fmul_rn((float)ay, 0.03125f);
    const float gx = __fsub_rn(1.f, fx), gy = __fsub_rn(1.f, fy);
    const double w0 = __fmul_rn(gy, gx), w1 = __fmul_rn(gy, fx), w2 = __fmul_rn(fy, gx), w3 = __fmul_rn(fy, fx);
    return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(v0, w0), __dmul_rn(v1, w1)), __dmul_rn(v2, w2)), __dmul_rn(v3, w3));
  }
};
template <> struct WarpAcc<unsigned char> {
  static __device__ __forceinline__ unsigned char run(int v0, int v1, int v2, int v3, int ax, int ay) {
    // 15-bit fixed-point weights: (32-ay)(32-ax)/1024 * 32768, exact
    const int w0 = (32 - ay) * (32 - ax) * 32, w1 = (32 - ay) * ax * 32, w2 = ay * (32 - ax) * 32, w3 = ay * ax * 32;
    const int s = (v0 * w0 + v1 * w1 + v2 * w2 + v3 * w3 + (1 << 14)) >> 15;
    return (unsigned char)min(max(s, 0), 255);
  }
};

template <typename T>
__global__ void __launch_bounds__(256) warp_affine_kernel(const T* __restrict__ src, int sH, int sW,
                                                          const double* __restrict__ matrices, T* __restrict__ dst,
                                                          int dH, int dW) {
  __shared__ double Mi[6];
  const int b = blockIdx.z;
  if (threadIdx.x == 0 && threadIdx.y == 0) {   // cv::warpAffine inverts the forward matrix in doubles
    const double* Mf = matrices + (size_t)b * 6;
    double m0 = Mf[0], m1 = Mf[1], m2 = Mf[2], m3 = Mf[3], m4 = Mf[4], m5 = Mf[5];
    double D = __dsub_rn(__dmul_rn(m0, m4), __dmul_rn(m1, m3));
    D = D != 0 ? __ddiv_rn(1.0, D) : 0.0;
    const double A11 = __dmul_rn(m4, D), A22 = __dmul_rn(m0, D);
    m0 = A11; m1 = __dmul_rn(m1, -D); m3 = __dmul_rn(m3, -D); m4 = A22;
    const double b1 = __dsub_rn(__dmul_rn(-m0, m2), __dmul_rn(m1, m5));
    const double b2 = __dsub_rn(__dmul_rn(-m3, m2), __dmul_rn(m4, m5));
    Mi[0] = m0; Mi[1] = m1; Mi[2] = b1; Mi[3] = m3; Mi[4] = m4; Mi[5] = b2;
  }
  __syncthreads();
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (x >= dW || y >= dH) return;
  src += (size_t)b * sH * sW;
  dst += (size_t)b * dH * dW;
  const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(Mi[0], (double)x), 1024.0));
  const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(Mi[3], (double)x), 1024.0));
  const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(Mi[1], (double)y), Mi[2]), 1024.0)) + 16;
  const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(Mi[4], (double)y), Mi[5]), 1024.0)) + 16;
  const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  const int sx = min(max(X >> 5, -32768), 32767), sy = min(max(Y >> 5, -32768), 32767);
  const int ax = X & 31, ay = Y & 31;
  const bool x0 = sx >= 0 && sx < sW, x1 = sx + 1 >= 0 && sx + 1 < sW;
  const bool y0 = sy >= 0 && sy < sH, y1 = sy + 1 >= 0 && sy + 1 < sH;
  const T zero = (T)0;
  const T v0 = (y0 && x0) ? src[(size_t)sy * sW + sx] : zero;
  const T v1 = (y0 && x1) ? src[(size_t)sy * sW + sx + 1] : zero;
  const T v2 = (y1 && x0) ? src[(size_t)(sy + 1) * sW + sx] : zero;
  const T v3 = (y1 && x1) ? src[(size_t)(sy + 1) * sW + sx + 1] : zero;
  dst[(size_t)y * dW + x] = WarpAcc<T>::run(v0, v1, v2, v3, ax, ay);
}

}  // namespace

size_t ransac_smem_bytes(int N) { return (size_t)N * (2 * sizeof(float2) + sizeof(int)); }

bool launch_ransac_affine_partial(LaunchCtx& ctx, const float* kpts0, const float* kpts1, const long long* matches0,
                                  const int* counts0, int B, int N, int M, double thr, int max_iters, double confidence,
                                  int refine, double* matrices, unsigned char* inlier0, int* info) {
  const size_t smem = ransac_smem_bytes(N);
  if (smem > 200 * 1024) return false;
  if (B <= 0) return true;
  static SmemOptIn opt;
  opt.ensure(ransac_affine_partial_kernel, 200 * 1024);
  {
    ProfScope ps(ctx, "ransac_affine_partial");
    ransac_affine_partial_kernel<<<B, kRansacThreads, smem, ctx.stream>>>(
        reinterpret_cast<const float2*>(kpts0), reinterpret_cast<const float2*>(kpts1), matches0, counts0, N, M, thr,
        max_iters, confidence, refine, matrices, inlier0, info);
  }
  B200M_LAUNCH_CHECK(ctx, "ransac_affine_partial");
  return true;
}

bool launch_warp_affine(LaunchCtx& ctx, const void* src, int dtype, int B, int sH, int sW, const double* matrices,
                        void* dst, int dH, int dW) {
  if (B <= 0 || dH <= 0 || dW <= 0) return true;
  dim3 grid(cdiv(dW, 32), cdiv(dH, 8), B), block(32, 8);
  {
    ProfScope ps(ctx, "warp_affine");
    if (dtype == 0)
      warp_affine_kernel<unsigned char><<<grid, block, 0, ctx.stream>>>((const unsigned char*)src, sH, sW, matrices,
                                                                        (unsigned char*)dst, dH, dW);
    else if (dtype == 1)
      warp_affine_kernel<float><<<grid, block, 0, ctx.stream>>>((const float*)src, sH, sW, matrices, (float*)dst, dH,
                                                                dW);
    else if (dtype == 2)
      warp_affine_kernel<double><<<grid, block, 0, ctx.stream>>>((const double*)src, sH, sW, matrices, (double*)dst,
                                                                 dH, dW);
    else
      return false;
  }
  B200M_LAUNCH_CHECK(ctx, "warp_affine");
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// cv2.resize(src, (dw, dh)) of the reference's data loader (datasets/SSHIDataset.py:20-22; uint8, INTER_LINEAR), restated
// from OpenCV 4.13 imgproc/src/resize.cpp: destination coordinate in double cast to float, 11-bit fixed-point weights
// (round half to even), horizontal taps clamped with their weight reset, vertical rows clipped with the weights kept,
// and OpenCV's substitution of its 2x2 box filter for an exact 2x decimation.  One thread per destination pixel; a warp
// writes 32 consecutive bytes and reads two source rows.  HBM-bound: (source bytes touched + destination bytes) / time.
struct ResizeCoef { int i0, i1, w0, w1; };
__device__ __forceinline__ ResizeCoef resize_coef(int d, int sn, double scale, bool vertical) {
  float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);      // no FMA contraction: OpenCV rounds twice
  int s = (int)floorf(f);
  f -= (float)s;
  if (!vertical) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= sn - 1) { f = 0.f; s = sn - 1; }
  }
  ResizeCoef c;
  c.w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  c.w1 = __float2int_rn(__fmul_rn(f, 2048.f));
  c.i0 = min(max(s, 0), sn - 1);
  c.i1 = min(max(s + 1, 0), sn - 1);
  return c;
}
__global__ void __launch_bounds__(256) resize_linear_u8_kernel(const unsigned char* __restrict__ src, int sh, int sw,
                                                               unsigned char* __restrict__ dst, int dh, int dw,
                                                               double scale_x, double scale_y, int area2) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= dw || y >= dh) return;
  const unsigned char* s = src + (size_t)blockIdx.z * sh * sw;
  int v;
  if (area2) {
    const unsigned char* r0 = s + (size_t)(2 * y) * sw + 2 * x;
    v = (r0[0] + r0[1] + r0[sw] + r0[sw + 1] + 2) >> 2;
  } else {
    const ResizeCoef cx = resize_coef(x, sw, scale_x, false), cy = resize_coef(y, sh, scale_y, true);
    const unsigned char* r0 = s + (size_t)cy.i0 * sw;
    const unsigned char* r1 = s + (size_t)cy.i1 * sw;
    const int h0 = r0[cx.i0] * cx.w0 + r0[cx.i1] * cx.w1;
    const int h1 = r1[cx.i0] * cx.w0 + r1[cx.i1] * cx.w1;
    v = (((cy.w0 * (h0 >> 4)) >> 16) + ((cy.w1 * (h1 >> 4)) >> 16) + 2) >> 2;
    v = min(max(v, 0), 255);
  }
  dst[((size_t)blockIdx.z * dh + y) * dw + x] = (unsigned char)v;
}

void launch_resize_linear_u8(LaunchCtx& ctx, const unsigned char* src, int B, int sh, int sw, unsigned char* dst, int dh,
                             int dw) {
  if (B <= 0 || dh <= 0 || dw <= 0) return;
  ProfScope prof__(ctx, "resize_u8");
  const double inv_x = (double)dw / sw, inv_y = (double)dh / sh;
  const double scale_x = 1.0 / inv_x, scale_y = 1.0 / inv_y;
  const int area2 = (int)scale_x == 2 && (int)scale_y == 2 && fabs(inv_x - 0.5) < 2.220446049250313e-16 &&
                    fabs(inv_y - 0.5) < 2.220446049250313e-16;
  dim3 grid(cdiv(dw, 32), cdiv(dh, 8), B);
  resize_linear_u8_kernel<<<grid, 256, 0, ctx.stream>>>(src, sh, sw, dst, dh, dw, scale_x, scale_y, area2);
  B200M_LAUNCH_CHECK(ctx, "resize_u8");
}

}  // namespace b200m
