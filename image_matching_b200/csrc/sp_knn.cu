// Descriptor nearest-neighbour matching with Lowe's ratio test: the step that follows SuperPoint in the reference's
// superpoint_flann_test.py:69-78 (cv2.FlannBasedMatcher KD-tree knnMatch(k=2) + `m.distance < 0.7 * n.distance`).
// Here the 2-NN search is EXACT (brute force, fp32 squared differences summed in channel order), which the
// approximate KD-tree search converges to; distances are Euclidean (not squared), like cv2's L2 matcher.
// Layout: descriptors (B, D, N) channel-major, as SuperPoint returns them.
//   block = 8 warps, each warp owns 8 query descriptors (in shared memory, broadcast reads) and sweeps the train set
//   32 descriptors at a time (lane = train descriptor: coalesced 128-byte rows of the (D, M) matrix, reused for the 8
//   queries from registers); per-lane best / second best, then a warp-level merge.
#include "kernels.cuh"

namespace b200m {


__device__ __forceinline__ void top2_push(float& d1, int& i1, float& d2, float d, int i) {
  if (d < d1 || (d == d1 && i < i1)) { d2 = d1; d1 = d; i1 = i; }
  else if (d < d2) d2 = d;
}

template <int D, int kKnnQ>      // kKnnQ = queries per warp (8 warps x kKnnQ x D floats of shared memory)
__global__ void __launch_bounds__(256) knn_ratio_kernel(const float* __restrict__ desc0, const float* __restrict__ desc1,
                                                        const int* __restrict__ counts0, const int* __restrict__ counts1,
                                                        int N, int M, float ratio, long long* __restrict__ match,
                                                        float* __restrict__ dist1, float* __restrict__ dist2) {
  __shared__ float q[8][kKnnQ][D];
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = counts0 ? counts0[b] : N, m = counts1 ? counts1[b] : M;
  const int i0 = (blockIdx.x * 8 + warp) * kKnnQ;
  const float* A = desc0 + (size_t)b * D * N;
  const float* Bm = desc1 + (size_t)b * D * M;
  for (int t = lane; t < kKnnQ * D; t += 32) {
    const int qi = t % kKnnQ, d = t / kKnnQ;
    q[warp][qi][d] = i0 + qi < N ? A[(size_t)d * N + i0 + qi] : 0.f;
  }
  __syncwarp();
  float d1[kKnnQ], d2[kKnnQ];
  int i1[kKnnQ];
#pragma unroll
  for (int k = 0; k < kKnnQ; ++k) { d1[k] = INFINITY; d2[k] = INFINITY; i1[k] = 0x7fffffff; }
  for (int j0 = 0; j0 < m; j0 += 32) {
    const int j = j0 + lane;
    float acc[kKnnQ];
#pragma unroll
    for (int k = 0; k < kKnnQ; ++k) acc[k] = 0.f;
    if (j < m) {
#pragma unroll 4
      for (int d = 0; d < D; ++d) {
        const float bv = __ldg(Bm + (size_t)d * M + j);
#pragma unroll
        for (int k = 0; k < kKnnQ; ++k) {
          const float df = q[warp][k][d] - bv;
          acc[k] = fmaf(df, df, acc[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < kKnnQ; ++k) top2_push(d1[k], i1[k], d2[k], acc[k], j);
    }
  }
  // merge the 32 lanes' (best, second) pairs
#pragma unroll
  for (int k = 0; k < kKnnQ; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od1 = __shfl_xor_sync(0xffffffffu, d1[k], o), od2 = __shfl_xor_sync(0xffffffffu, d2[k], o);
      const int oi1 = __shfl_xor_sync(0xffffffffu, i1[k], o);
      if (od1 < d1[k] || (od1 == d1[k] && oi1 < i1[k])) {
        d2[k] = fminf(d1[k], od2);
        d1[k] = od1; i1[k] = oi1;
      } else {
        d2[k] = fminf(d2[k], od1);
      }
    }
    const int i = i0 + k;
    if (lane == 0 && i < N) {
      const float e1 = sqrtf(d1[k]), e2 = sqrtf(d2[k]);
      const bool valid = i < n && m >= 2 && e1 < ratio * e2;          // superpoint_flann_test.py:76-78
      match[(size_t)b * N + i] = valid ? (long long)i1[k] : -1;
      dist1[(size_t)b * N + i] = i < n && m >= 1 ? e1 : 0.f;
      dist2[(size_t)b * N + i] = i < n && m >= 2 ? e2 : 0.f;
    }
  }
}

bool launch_knn_ratio(LaunchCtx& ctx, const float* desc0, const float* desc1, const int* counts0, const int* counts1,
                      int B, int D, int N, int M, float ratio, long long* match, float* dist1, float* dist2) {
  if (B <= 0 || N <= 0) return true;
  ProfScope prof__(ctx, "knn_ratio");
  if (D == 64) knn_ratio_kernel<64, 8><<<dim3(cdiv(N, 64), B), 256, 0, ctx.stream>>>(desc0, desc1, counts0, counts1, N, M, ratio, match, dist1, dist2);
  else if (D == 128) knn_ratio_kernel<128, 8><<<dim3(cdiv(N, 64), B), 256, 0, ctx.stream>>>(desc0, desc1, counts0, counts1, N, M, ratio, match, dist1, dist2);
  else if (D == 256) knn_ratio_kernel<256, 4><<<dim3(cdiv(N, 32), B), 256, 0, ctx.stream>>>(desc0, desc1, counts0, counts1, N, M, ratio, match, dist1, dist2);
  else return false;
  B200M_LAUNCH_CHECK(ctx, "knn_ratio");
  return true;
}

}  // namespace b200m
