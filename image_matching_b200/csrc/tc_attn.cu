// SuperGlue multi-head attention on tcgen05 tensor cores: flash-style (online softmax, the (B,4,N,M)
// probability tensor of the reference is never materialised), fp32-class accuracy via 3xTF32.
// Reference: superglue/models/superglue_test.py:85-89 (attention), :92-107 (MultiHeadedAttention).
//
// One CTA per (side, pair, head, 128-query tile); key tiles of KT keys stream through a 2-stage ring.
//   warp 0      TMA producer : Q tile once, then K / V^T tiles (hi and lo planes): 2-D tensor maps with 128-byte
//                              swizzle, box = 128 B x rows, i.e. the canonical K-major SWIZZLE_128B UMMA layout.
//                              K tiles are [keys][32 ch]; V is read from the transposed copy V^T [ch][keys] that the
//                              q|k|v projection's epilogue writes next to it, so both MMAs take K-major operands
//                              (kind::tf32 returned zeros for an MN-major B operand on this part).
//   warp 1      MMA issuer   : S_j = Q K_j^T (M=128, N=KT, K=d) into a double-buffered TMEM tile, then
//                              OT_j = P_j V_j (M=128, N=d, K=KT); each product is hi*hi + hi*lo + lo*hi.
//   warps 2..5  softmax      : thread = query row.  tcgen05.ld S_j, running max / sum (exp2 with the 1/sqrt(d)
//                              scale folded into one FFMA), P_j split into tf32 hi/lo and stored to shared memory
//                              as the next MMA's A operand, OT_{j-1} pulled from TMEM and folded into the
//                              register accumulator with the usual exp(m_old - m_new) correction.
#include "kernels.cuh"
#include "tc_common.cuh"

namespace b200m {

using namespace tc;

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, float* v) {
  if constexpr (N == 16) {
    tmem_ld16(taddr, v);
  } else {
#pragma unroll
    for (int c = 0; c < N / 32; ++c) tmem_ld32(taddr + c * 32, v + c * 32);
  }
}

constexpr int kTaQ = 128;   // queries per CTA

template <int HD, int KT>
struct TcAttnSmem {
  static constexpr int Q_PLANE = kTaQ * HD * 4;          // bytes, one (hi or lo) plane
  static constexpr int KV_PLANE = KT * HD * 4;
  static constexpr int KV_STAGE = 4 * KV_PLANE;          // K hi, K lo, V hi, V lo
  static constexpr int P_PLANE = kTaQ * KT * 4;
  static constexpr int OFF_KV = 2 * Q_PLANE;
  static constexpr int NKV = 3;                           // K/V ring depth (TMA latency ~ 2-3 tiles of work)
  static constexpr int OFF_P = OFF_KV + NKV * KV_STAGE;
  static constexpr int OFF_BAR = OFF_P + 2 * P_PLANE;
  static constexpr int N_BARS = 1 + 3 + 3 + 2 + 2 + 1 + 1 + 2 + 2;
  static constexpr size_t BYTES = 1024 + OFF_BAR + N_BARS * 8 + 16;   // 1024: SWIZZLE_128B tiles need 1 KB alignment
  static constexpr int TMEM_COLS = (2 * KT + 2 * HD) <= 128 ? 128 : 256;
  static constexpr int NH = HD / 32;                      // 128-byte column halves per row
  static constexpr int Q_HALF = kTaQ * 128;               // bytes of one 32-channel half of a Q plane
  static constexpr int KV_HALF = KT * 128;
  static constexpr int VT_CHUNK = HD * 128;               // bytes of one 32-key chunk of a V^T plane
};

struct TcAttnParams {
  float* msg;              // [rows][D] head-major message (pre-merge)
  int B, Np, D;
  const int* counts0; const int* counts1;
  int n_full0, n_full1;
  int cross;
  float scale_log2e;       // log2(e) / sqrt(d)
};

template <int HD, int KT>
__global__ void __launch_bounds__(192, 1)
tc_attention_kernel(const __grid_constant__ CUtensorMap tm_q_hi, const __grid_constant__ CUtensorMap tm_q_lo,
                    const __grid_constant__ CUtensorMap tm_kv_hi, const __grid_constant__ CUtensorMap tm_kv_lo,
                    const __grid_constant__ CUtensorMap tm_vt_hi, const __grid_constant__ CUtensorMap tm_vt_lo,
                    TcAttnParams p) {
  using SM = TcAttnSmem<HD, KT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + SM::OFF_KV;
  uint8_t* sP = smem + SM::OFF_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* kv_full = q_full + 1;
  uint64_t* kv_empty = kv_full + SM::NKV;
  uint64_t* s_full = kv_empty + SM::NKV;
  uint64_t* s_empty = s_full + 2;
  uint64_t* p_full = s_empty + 2;
  uint64_t* p_empty = p_full + 1;
  uint64_t* o_full = p_empty + 1;
  uint64_t* o_empty = o_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int side = blockIdx.z / p.B, b = blockIdx.z - side * p.B;
  const int src = p.cross ? 1 - side : side;
  const int n_k = src == 0 ? (p.counts0 ? p.counts0[b] : p.n_full0) : (p.counts1 ? p.counts1[b] : p.n_full1);
  const int T = cdiv(n_k, KT);
  const int q_row0 = (side * p.B + b) * p.Np + blockIdx.x * kTaQ;     // global row of the first query
  const int k_row0 = (src * p.B + b) * p.Np;
  const int cq = head * HD, ck = p.D + head * HD;          // first column of this head's q / k
  const int vt_row0 = (src * p.B + b) * p.D + head * HD;   // first row of this head in V^T [block][D][Np]

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < SM::NKV; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4);
      mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 4);
    }
    mbar_init(p_full, 4);
    mbar_init(p_empty, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_q_hi); tma_prefetch_desc(&tm_q_lo);
    tma_prefetch_desc(&tm_kv_hi); tma_prefetch_desc(&tm_kv_lo);
    tma_prefetch_desc(&tm_vt_hi); tma_prefetch_desc(&tm_vt_lo);
  }
  if (warp == 1) tmem_alloc(tmem_slot, SM::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 2 * KT;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (T > 0) {
      mbar_expect_tx(q_full, 2 * SM::Q_PLANE);
#pragma unroll
      for (int hf = 0; hf < SM::NH; ++hf) {
        tma_load_2d(sQ + hf * SM::Q_HALF, &tm_q_hi, q_full, cq + hf * 32, q_row0);
        tma_load_2d(sQ + SM::Q_PLANE + hf * SM::Q_HALF, &tm_q_lo, q_full, cq + hf * 32, q_row0);
      }
    }
    for (int j = 0; j < T; ++j) {
      const int st = j % SM::NKV, ph = (j / SM::NKV) & 1;
      mbar_wait(&kv_empty[st], ph ^ 1);
      mbar_expect_tx(&kv_full[st], SM::KV_STAGE);
      uint8_t* dst = sKV + st * SM::KV_STAGE;
      const int r = k_row0 + j * KT;
#pragma unroll
      for (int hf = 0; hf < SM::NH; ++hf) {
        tma_load_2d(dst + hf * SM::KV_HALF, &tm_kv_hi, &kv_full[st], ck + hf * 32, r);
        tma_load_2d(dst + SM::KV_PLANE + hf * SM::KV_HALF, &tm_kv_lo, &kv_full[st], ck + hf * 32, r);
      }
#pragma unroll
      for (int kc = 0; kc < KT / 32; ++kc) {   // V^T: [HD channel rows][32 keys] per chunk
        tma_load_2d(dst + 2 * SM::KV_PLANE + kc * SM::VT_CHUNK, &tm_vt_hi, &kv_full[st], j * KT + kc * 32, vt_row0);
        tma_load_2d(dst + 3 * SM::KV_PLANE + kc * SM::VT_CHUNK, &tm_vt_lo, &kv_full[st], j * KT + kc * 32, vt_row0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (converged warp)
    if (T > 0) {
      const uint32_t idesc_s = instr_desc(2, 128, KT);                    // A, B K-major
      const uint32_t idesc_o = instr_desc(2, 128, HD);                    // A = P, B = V^T, both K-major
      const uint32_t q_base = smem_u32(sQ), p_base = smem_u32(sP);
      auto issue_S = [&](int j) {
        const int st = j & 1;
        const uint32_t k_base = smem_u32(sKV + (j % SM::NKV) * SM::KV_STAGE);
        if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < HD / 8; ++ks) {
          // K-major SWIZZLE_128B: 8-channel step = +32 B inside the 128 B row, next 32 channels = next half
          const uint32_t qo = (ks >> 2) * SM::Q_HALF + (ks & 3) * 32, ko = (ks >> 2) * SM::KV_HALF + (ks & 3) * 32;
          const uint64_t qh = smem_desc_sw128(q_base + qo);
          const uint64_t ql = smem_desc_sw128(q_base + SM::Q_PLANE + qo);
          const uint64_t kh = smem_desc_sw128(k_base + ko);
          const uint64_t kl = smem_desc_sw128(k_base + SM::KV_PLANE + ko);
          mma_tf32(tS + st * KT, qh, kh, idesc_s, ks != 0);
          mma_tf32(tS + st * KT, qh, kl, idesc_s, 1);
          mma_tf32(tS + st * KT, ql, kh, idesc_s, 1);
        }
        tc_commit(&s_full[st]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_S(0);
      for (int j = 0; j < T; ++j) {
        const int st = j & 1, ph = (j >> 1) & 1;
        if (j + 1 < T) {
          const int sn = (j + 1) & 1, pn = ((j + 1) >> 1) & 1;
          mbar_wait(&kv_full[(j + 1) % SM::NKV], ((j + 1) / SM::NKV) & 1);
          mbar_wait(&s_empty[sn], pn ^ 1);
          tc_fence_after();
          issue_S(j + 1);
        }
        mbar_wait(p_full, j & 1);
        mbar_wait(&o_empty[st], ph ^ 1);
        tc_fence_after();
        const uint32_t v_base = smem_u32(sKV + (j % SM::NKV) * SM::KV_STAGE + 2 * SM::KV_PLANE);
        if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < KT / 8; ++ks) {
          const uint64_t ph_ = smem_desc_nosw(p_base + ks * 2 * (kTaQ * 16), kTaQ * 16, 128);
          const uint64_t pl_ = smem_desc_nosw(p_base + SM::P_PLANE + ks * 2 * (kTaQ * 16), kTaQ * 16, 128);
          // V^T, K-major SWIZZLE_128B: rows = channels, 32 keys (128 B) per row; 8-key step = +32 B
          const uint32_t vo = (ks >> 2) * SM::VT_CHUNK + (ks & 3) * 32;
          const uint64_t vh = smem_desc_sw128(v_base + vo);
          const uint64_t vl = smem_desc_sw128(v_base + SM::KV_PLANE + vo);
          mma_tf32(tO + st * HD, ph_, vh, idesc_o, ks != 0);
          mma_tf32(tO + st * HD, ph_, vl, idesc_o, 1);
          mma_tf32(tO + st * HD, pl_, vh, idesc_o, 1);
        }
        tc_commit(&kv_empty[j % SM::NKV]);
        tc_commit(p_empty);
        tc_commit(&o_full[st]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------------ softmax / accumulate (thread = query row)
    const int w4 = warp & 3;
    const int m = w4 * 32 + lane;
    const uint32_t lane_base = (uint32_t)(w4 * 32) << 16;
    const float c = p.scale_log2e;
    float o[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, m_o = -INFINITY, m_prev = -INFINITY;
    float4* Ph = reinterpret_cast<float4*>(sP) + m;                       // [key chunk][row][4]
    float4* Pl = reinterpret_cast<float4*>(sP + SM::P_PLANE) + m;
    auto fold_O = [&](int j, float m_tile) {
      const int st = j & 1, ph = (j >> 1) & 1;
      mbar_wait(&o_full[st], ph);
      tc_fence_after();
      float ot[HD];
      tmem_ld_n<HD>(tO + lane_base + st * HD, ot);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[st]);
      const float corr = exp2f((m_o - m_tile) * c);                      // exp2(-inf) = 0 on the first tile
#pragma unroll
      for (int i = 0; i < HD; ++i) o[i] = fmaf(o[i], corr, ot[i]);
      m_o = m_tile;
    };
    for (int j = 0; j < T; ++j) {
      const int st = j & 1, ph = (j >> 1) & 1;
      mbar_wait(&s_full[st], ph);
      tc_fence_after();
      float s[KT];
      tmem_ld_n<KT>(tS + lane_base + st * KT, s);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[st]);
      const int kbase = j * KT;
      if (kbase + KT > n_k) {
#pragma unroll
        for (int i = 0; i < KT; ++i)
          if (kbase + i >= n_k) s[i] = -INFINITY;
      }
      float mx = s[0];
#pragma unroll
      for (int i = 1; i < KT; ++i) mx = fmaxf(mx, s[i]);
      const float m_new = fmaxf(m_run, mx);
      const float neg = -m_new * c;
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < KT; ++i) {
        s[i] = exp2f(fmaf(s[i], c, neg));
        sum += s[i];
      }
      l_run = l_run * exp2f((m_run - m_new) * c) + sum;
      m_run = m_new;
      // P_j -> shared memory (A operand of the PV MMA), tf32 hi / lo planes
      mbar_wait(p_empty, (j & 1) ^ 1);
#pragma unroll
      for (int g = 0; g < KT / 4; ++g) {
        float h[4], l[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          h[t] = __uint_as_float((__float_as_uint(s[4 * g + t]) + 0x1000u) & 0xFFFFE000u);
          l[t] = s[4 * g + t] - h[t];
        }
        Ph[g * kTaQ] = make_float4(h[0], h[1], h[2], h[3]);
        Pl[g * kTaQ] = make_float4(l[0], l[1], l[2], l[3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (j > 0) fold_O(j - 1, m_prev);
      m_prev = m_new;
    }
    if (T > 0) fold_O(T - 1, m_prev);
    const int row = blockIdx.x * kTaQ + m;
    if (row < p.Np) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      float4* dst = reinterpret_cast<float4*>(p.msg + ((size_t)(side * p.B + b) * p.Np + row) * p.D + head * HD);
#pragma unroll
      for (int g = 0; g < HD / 4; ++g)
        dst[g] = make_float4(o[4 * g] * inv, o[4 * g + 1] * inv, o[4 * g + 2] * inv, o[4 * g + 3] * inv);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, SM::TMEM_COLS);
  }
}

// 2-D view of a [rows][ld] fp32 plane; box = 32 columns (128 B) x box_rows, 128-byte swizzle
static bool make_sw128_map(CUtensorMap* m, const float* base, size_t rows, int ld, int box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int HD, int KT>
static bool launch_tc_attn_t(LaunchCtx& ctx, const float* qkv_hi, const float* qkv_lo, const float* vt_hi,
                             const float* vt_lo, float* msg, int B, int Np,
                             int D, int heads, const int* c0, const int* c1, int nf0, int nf1, bool cross) {
  ProfScope prof__(ctx, "tc_attention");
  const size_t rows = (size_t)2 * B * Np;
  CUtensorMap mq_hi, mq_lo, mk_hi, mk_lo, mv_hi, mv_lo;
  if (!make_sw128_map(&mq_hi, qkv_hi, rows, 3 * D, kTaQ) || !make_sw128_map(&mq_lo, qkv_lo, rows, 3 * D, kTaQ) ||
      !make_sw128_map(&mk_hi, qkv_hi, rows, 3 * D, KT) || !make_sw128_map(&mk_lo, qkv_lo, rows, 3 * D, KT) ||
      !make_sw128_map(&mv_hi, vt_hi, (size_t)2 * B * D, Np, HD) || !make_sw128_map(&mv_lo, vt_lo, (size_t)2 * B * D, Np, HD))
    return false;
  using SM = TcAttnSmem<HD, KT>;
  static bool attr_set = false;
  auto kern = tc_attention_kernel<HD, KT>;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::BYTES) != cudaSuccess) return false;
    attr_set = true;
  }
  TcAttnParams p;
  p.msg = msg; p.B = B; p.Np = Np; p.D = D; p.counts0 = c0; p.counts1 = c1; p.n_full0 = nf0; p.n_full1 = nf1;
  p.cross = cross ? 1 : 0;
  p.scale_log2e = 1.4426950408889634f / sqrtf((float)HD);
  dim3 grid(cdiv(Np, kTaQ), heads, 2 * B);
  kern<<<grid, 192, SM::BYTES, ctx.stream>>>(mq_hi, mq_lo, mk_hi, mk_lo, mv_hi, mv_lo, p);
  B200M_LAUNCH_CHECK(ctx, "tc_attention");
  return true;
}

bool launch_tc_attention(LaunchCtx& ctx, const float* qkv_hi, const float* qkv_lo, const float* vt_hi,
                         const float* vt_lo, float* msg, int B, int Np, int D,
                         int heads, const int* counts0, const int* counts1, int n_full0, int n_full1, bool cross) {
  const int hd = D / heads;
  // hd = 16 (D = 64) rows are 64 B -- would need the 64-byte swizzle variant; the fp32 CUDA-core kernel handles it
  if (hd == 32) return launch_tc_attn_t<32, 64>(ctx, qkv_hi, qkv_lo, vt_hi, vt_lo, msg, B, Np, D, heads, counts0, counts1, n_full0, n_full1, cross);
  if (hd == 64) return launch_tc_attn_t<64, 32>(ctx, qkv_hi, qkv_lo, vt_hi, vt_lo, msg, B, Np, D, heads, counts0, counts1, n_full0, n_full1, cross);
  return false;
}

}  // namespace b200m
