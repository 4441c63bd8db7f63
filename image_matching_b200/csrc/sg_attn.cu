// SuperGlue multi-head attention, flash-style (the (B,4,N,M) probability tensor of the reference is
// never materialised), fp32 CUDA-core path.
// Reference: superglue/models/superglue_test.py:85-89 (attention), :92-107 (MultiHeadedAttention).
// Heads are de-interleaved at pack time (reference channel c = dd*4 + h  ->  column h*d + dd), so a head's
// q/k/v slice is a contiguous run of d floats in a token row.
#include "kernels.cuh"

namespace b200m {

constexpr int kAttQ = 64;    // queries per block (4 warps x 16)
constexpr int kAttK = 64;    // keys per tile
constexpr int kPsLd = 20;    // padded row (16 probabilities) -> conflict-free float4 stores

template <int HD>
struct AttnSmem {
  static constexpr int LDK = kAttK + 1;
  static constexpr int Q_F = kAttQ * HD;
  static constexpr int KT_F = ((HD * LDK + 3) / 4) * 4;
  static constexpr int V_F = kAttK * HD;
  static constexpr int PS_F = 4 * kAttK * kPsLd;
  static constexpr size_t BYTES = (size_t)(Q_F + KT_F + V_F + PS_F) * sizeof(float);
};

template <int HD>
__global__ void __launch_bounds__(128) attention_kernel(const float* __restrict__ qkv, float* __restrict__ msg,
                                                        int B, int Np, int D,
                                                        const int* __restrict__ counts0,
                                                        const int* __restrict__ counts1, int n_full0, int n_full1,
                                                        int cross, float scale) {
  using SM = AttnSmem<HD>;
  constexpr int LDK = SM::LDK;
  constexpr int OD = (HD + 31) / 32;   // output dims owned per lane
  extern __shared__ float4 smem4[];
  float* Qs = reinterpret_cast<float*>(smem4);
  float* Kt = Qs + SM::Q_F;
  float* Vs = Kt + SM::KT_F;
  float* Ps = Vs + SM::V_F;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int head = blockIdx.y;
  const int side = blockIdx.z / B, b = blockIdx.z - side * B;
  const int src = cross ? 1 - side : side;
  const int n_k = src == 0 ? (counts0 ? counts0[b] : n_full0) : (counts1 ? counts1[b] : n_full1);
  const int ld = 3 * D;
  const int q0 = blockIdx.x * kAttQ;
  const float* qbase = qkv + ((size_t)(side * B + b) * Np) * ld + head * HD;
  const float* kbase = qkv + ((size_t)(src * B + b) * Np) * ld + D + head * HD;
  const float* vbase = kbase + D;
  constexpr int F4 = HD / 4;

  for (int i = tid; i < kAttQ * F4; i += 128) {
    int r = i / F4, c = i - r * F4;
    int row = q0 + r;
    float4 v = row < Np ? *reinterpret_cast<const float4*>(qbase + (size_t)row * ld + c * 4) : make_float4(0, 0, 0, 0);
    reinterpret_cast<float4*>(Qs)[i] = v;
  }

  float m_run[16], l_run[16], o[16][OD];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
#pragma unroll
    for (int t = 0; t < OD; ++t) o[i][t] = 0.f;
  }
  float* Pw = Ps + w * kAttK * kPsLd;
  const float* Qw = Qs + w * 16 * HD;

  const int ntiles = cdiv(n_k, kAttK);
  for (int kt = 0; kt < ntiles; ++kt) {
    const int k0 = kt * kAttK;
    __syncthreads();   // previous tile fully consumed (also orders the Q stores on the first pass)
    for (int i = tid; i < kAttK * F4; i += 128) {
      int r = i / F4, c = i - r * F4;
      int row = k0 + r;
      float4 kv = make_float4(0, 0, 0, 0), vv = kv;
      if (row < Np) {
        kv = *reinterpret_cast<const float4*>(kbase + (size_t)row * ld + c * 4);
        vv = *reinterpret_cast<const float4*>(vbase + (size_t)row * ld + c * 4);
      }
      Kt[(c * 4 + 0) * LDK + r] = kv.x; Kt[(c * 4 + 1) * LDK + r] = kv.y;
      Kt[(c * 4 + 2) * LDK + r] = kv.z; Kt[(c * 4 + 3) * LDK + r] = kv.w;
      reinterpret_cast<float4*>(Vs)[i] = vv;
    }
    __syncthreads();

    // ---- S = Q K^T for 16 queries x 2 keys per lane
    float s[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i][0] = s[i][1] = 0.f;
#pragma unroll 2
    for (int d4 = 0; d4 < F4; ++d4) {
      float ka[4], kb[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        ka[c] = Kt[(d4 * 4 + c) * LDK + lane];
        kb[c] = Kt[(d4 * 4 + c) * LDK + lane + 32];
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float4 q = *reinterpret_cast<const float4*>(Qw + i * HD + d4 * 4);
        s[i][0] = fmaf(q.x, ka[0], s[i][0]); s[i][0] = fmaf(q.y, ka[1], s[i][0]);
        s[i][0] = fmaf(q.z, ka[2], s[i][0]); s[i][0] = fmaf(q.w, ka[3], s[i][0]);
        s[i][1] = fmaf(q.x, kb[0], s[i][1]); s[i][1] = fmaf(q.y, kb[1], s[i][1]);
        s[i][1] = fmaf(q.z, kb[2], s[i][1]); s[i][1] = fmaf(q.w, kb[3], s[i][1]);
      }
    }
    // ---- online softmax (fp32, exact expf)
    const bool v0 = (k0 + lane) < n_k, v1 = (k0 + lane + 32) < n_k;
    float corr[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float a = v0 ? s[i][0] * scale : -INFINITY;
      float c = v1 ? s[i][1] * scale : -INFINITY;
      float mt = warp_max(fmaxf(a, c));
      float mn = fmaxf(m_run[i], mt);
      float mref = (mn == -INFINITY) ? 0.f : mn;
      float p0 = expf(a - mref), p1 = expf(c - mref);
      float lt = warp_sum(p0 + p1);
      corr[i] = expf(m_run[i] - mref);
      l_run[i] = l_run[i] * corr[i] + lt;
      m_run[i] = mn;
      s[i][0] = p0;
      s[i][1] = p1;
    }
    __syncwarp();
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
      *reinterpret_cast<float4*>(Pw + lane * kPsLd + i4 * 4) =
          make_float4(s[i4 * 4][0], s[i4 * 4 + 1][0], s[i4 * 4 + 2][0], s[i4 * 4 + 3][0]);
      *reinterpret_cast<float4*>(Pw + (lane + 32) * kPsLd + i4 * 4) =
          make_float4(s[i4 * 4][1], s[i4 * 4 + 1][1], s[i4 * 4 + 2][1], s[i4 * 4 + 3][1]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i)
#pragma unroll
      for (int t = 0; t < OD; ++t) o[i][t] *= corr[i];
    __syncwarp();
    // ---- O += P V
#pragma unroll 4
    for (int j = 0; j < kAttK; ++j) {
      float vv[OD];
#pragma unroll
      for (int t = 0; t < OD; ++t) {
        int dim = lane + 32 * t;
        vv[t] = dim < HD ? Vs[j * HD + dim] : 0.f;
      }
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        float4 p = *reinterpret_cast<const float4*>(Pw + j * kPsLd + i4 * 4);
#pragma unroll
        for (int t = 0; t < OD; ++t) {
          o[i4 * 4 + 0][t] = fmaf(p.x, vv[t], o[i4 * 4 + 0][t]);
          o[i4 * 4 + 1][t] = fmaf(p.y, vv[t], o[i4 * 4 + 1][t]);
          o[i4 * 4 + 2][t] = fmaf(p.z, vv[t], o[i4 * 4 + 2][t]);
          o[i4 * 4 + 3][t] = fmaf(p.w, vv[t], o[i4 * 4 + 3][t]);
        }
      }
    }
    __syncwarp();
  }
  // ---- normalise and store (head-major message row)
  float* mbase = msg + ((size_t)(side * B + b) * Np) * D + head * HD;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    int row = q0 + w * 16 + i;
    if (row >= Np) continue;
    float inv = l_run[i] > 0.f ? 1.f / l_run[i] : 0.f;
#pragma unroll
    for (int t = 0; t < OD; ++t) {
      int dim = lane + 32 * t;
      if (dim < HD) mbase[(size_t)row * D + dim] = o[i][t] * inv;
    }
  }
}

template <int HD>
static void launch_attention_t(LaunchCtx& ctx, const float* qkv, float* msg, int B, int Np, int D, int heads,
                               const int* c0, const int* c1, int nf0, int nf1, bool cross) {
  ProfScope prof__(ctx, "attention");
  static SmemOptIn opt;
  auto kern = attention_kernel<HD>;
  opt.ensure(kern, (int)AttnSmem<HD>::BYTES);
  dim3 grid(cdiv(Np, kAttQ), heads, 2 * B);
  float scale = 1.f / sqrtf((float)HD);
  kern<<<grid, 128, AttnSmem<HD>::BYTES, ctx.stream>>>(qkv, msg, B, Np, D, c0, c1, nf0, nf1, cross ? 1 : 0, scale);
  B200M_LAUNCH_CHECK(ctx, "attention");
}

void launch_attention(LaunchCtx& ctx, const float* qkv, float* msg, int B, int Np, int D, int heads,
                      const int* counts0, const int* counts1, int n_full0, int n_full1, bool cross) {
  int hd = D / heads;
  if (hd == 16) launch_attention_t<16>(ctx, qkv, msg, B, Np, D, heads, counts0, counts1, n_full0, n_full1, cross);
  else if (hd == 32) launch_attention_t<32>(ctx, qkv, msg, B, Np, D, heads, counts0, counts1, n_full0, n_full1, cross);
  else launch_attention_t<64>(ctx, qkv, msg, B, Np, D, heads, counts0, counts1, n_full0, n_full1, cross);
}

}  // namespace b200m
