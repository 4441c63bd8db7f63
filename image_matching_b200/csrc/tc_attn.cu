// SuperGlue multi-head attention on tcgen05 tensor cores: flash-style (online softmax, the (B,4,N,M)
// probability tensor of the reference is never materialised), fp32-class accuracy via a 2-term fp16 split
// (x = hi + lo, hi = fp16(x), lo = fp16(x - hi); every product is hi*hi + hi*lo + lo*hi, exact in the fp32
// accumulator; |q|,|k|,|v| = O(1) and p in [0, 2^8], so the unscaled lo term costs < 3e-8 relative).
// Reference: superglue/models/superglue_test.py:85-89 (attention), :92-107 (MultiHeadedAttention).
// Instantiated for head_dim 16 / 32 / 64 (descriptor_dim 64 / 128 / 256: the widths of the reference's SuperPoint checkpoints).
//
// Round-2 rewrite.  The round-1 kernel spent ~12 instructions per score element in the softmax warps (issue-bound:
// 5.9 ms per 64-pair step against a 2.3 ms MUFU floor); this version needs ~4.5:
//   * P never goes through shared memory: the softmax threads write the fp16 hi / lo planes of P with tcgen05.st INTO THE
//     TMEM COLUMNS OF THE S TILE THEY JUST READ (a thread's 32 fp32 scores become 16 + 16 packed words), and the P.V MMA
//     takes its A operand from tensor memory.  No st.shared, no fence.proxy.async per tile, no p_empty / s_empty
//     barriers (the tensor pipe executes in issue order: Q.K^T of tile j+2 into the same columns is issued after P.V of
//     tile j), and the CTA shrinks by the 32 KB of P planes.
//   * d = 32: the row sum comes from the tensor core: the V^T tile carries 16 extra rows of ones, so P.V (N = d + 16) also
//     accumulates sum_j (p_hi + p_lo)_j -- exactly the weights that multiply V -- next to O.  No FADD per element.
//     d = 64: the threads keep the sums (the MUFU, not the FMA pipe, paces them there: two MMAs' worth of tensor work per
//     exponential) so that O needs only d columns and 64-key tiles fit tensor memory (-11 % against 32-key tiles).
//   * no running maximum: exponentials are taken against a reference that is the row maximum of the FIRST key tile and
//     only moves when a later p exceeds 2^8 (detected on the packed fp16 words with one HMNMX2 per two elements; rare).
//     Only then the scores are re-read, the reference is raised and O (with its sum columns) is rescaled in TMEM.
//   * the hi / lo split uses the sm_100 mixed-precision FMA (fma.rn.f32.f16: x - float(h) in ONE instruction taking the
//     packed half directly), 2 instructions per element instead of 3.
//
// One CTA per (side, pair, head, 128-query tile), two CTAs per SM; key tiles of KT = 64 keys stream through a TMA ring
// (4 stages at d = 32, 2 at d = 64).  Where the time goes at 1024 keys (clock64 stamps of all 4096 CTAs, B200M_ATTN_TRACE):
// a CTA lives ~27 k cycles -- 1.5 k set-up, 2.0 k until the first score tile, 20.8 k in the 16-tile loop, 2.2 k output, 0.8 k
// exit sync -- and its slot is re-filled 2.3 k cycles later.  The loop runs at the MUFU rate (two co-resident CTAs: 4 softmax
// warps per scheduler x 256 MUFU cycles per tile ~ 80 % of the ~1.3 k-cycle tile period); the fixed costs overlap the
// other CTA's loop, which alone is bound by its single MMA-issuing thread (a warp issues one tcgen05.mma per ~45 cycles
// whatever its size: 18 per tile).  Tried and measured worse: a persistent two-CTA-per-SM variant (the resident CTAs
// queue on the tensor pipe: +12 %, with or without a deliberate half-item offset between them), P_hi x [V_hi ; V_lo] as one
// N = 2d MMA with thread-side sums (+4 %), two MMA-issuing warps that take the even / odd key tiles (neutral at 1024
// keys, -1.6 % at 4096: the 80-register cap of 352 threads eats the gain).  Kept: Q in tensor memory (-7 %).
//   warp 0      TMA producer : Q tile once, then K / V^T tiles (hi and lo planes), 2-D tensor maps, 128- or 64-byte
//                              swizzle (= row length), i.e. the canonical K-major swizzled UMMA layouts.  K tiles are
//                              [keys][d]; V is read from the transposed copy V^T [d][keys] that the q|k|v projection's
//                              epilogue writes, so both MMAs take K-major B operands.
//   warp 1      MMA issuer   : S_j = Q K_j^T (M=128, N=KT, K=d; A, B from shared memory) into a double-buffered TMEM
//                              tile, then [O | l] += P_j [V_j ; 1] (M=128, N=d+16, K=KT; A = P from TMEM), one accumulator
//                              per key half.
//   warps 2..9  softmax      : thread = (query row, key half): tcgen05.ld 32 scores, p = ex2.approx((s - ref) c), split,
//                              tcgen05.st.  The two key halves keep independent references / accumulators (2-way
//                              split-KV) and are merged once at the end.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_fp16.h>
#include "kernels.cuh"
#include "tc_common.cuh"

namespace b200m {

using namespace tc;

__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// N consecutive fp32 columns of this thread's TMEM lane (N = 16, 32, 48, 64, 80: chunks of 32 then one of 16)
template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, float* v) {
  static_assert(N % 16 == 0, "16-column granularity");
#pragma unroll
  for (int c = 0; c < N / 32; ++c) tmem_ld32_issue(taddr + c * 32, v + c * 32);
  if constexpr (N % 32 == 16) tmem_ld16_issue(taddr + (N / 32) * 32, v + (N / 32) * 32);
  tmem_ld_wait();
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_st_words(uint32_t taddr, const uint32_t* v) {     // N 32-bit words, no wait
  static_assert(N == 8 || N == 16, "8 or 16 words");
  if constexpr (N == 16) tmem_st16(taddr, v); else tmem_st8(taddr, v);
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T : A (M = 128 rows = TMEM lanes, K fp16 values packed two per 32-bit column, even k in
// the low half) is read from tensor memory
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ex2.approx.ftz: 1 MUFU op, max relative error 2^-22 (the accurate exp2f expands to ~4 extra instructions)
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// K-major swizzled descriptor for rows of ROWB bytes (128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B, 32 -> SWIZZLE_32B); 8-row
// groups are 8*ROWB bytes apart
template <int ROWB>
__device__ __forceinline__ uint64_t smem_desc_sw(uint32_t saddr) {
  static_assert(ROWB == 128 || ROWB == 64 || ROWB == 32, "row length must be 32, 64 or 128 bytes");
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * ROWB) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(ROWB == 128 ? 2 : ROWB == 64 ? 4 : 6) << 61;
  return d;
}

constexpr int kTaQ = 128;   // queries per CTA

template <int HD, int KT>
struct TcAttnSmem {
  static constexpr int QROW = HD * 2;                     // bytes per Q / K row (fp16)
  static constexpr int VROW = KT * 2;                     // bytes per V^T row
  // P.V accumulator columns per key half.  d = 32: O (d) | 16 copies of the row sum -- the V^T hi tile carries 16 extra
  // rows of ones, so the tensor core also accumulates sum_j (p_hi + p_lo)_j, exactly the weights that multiply V, and the
  // softmax threads spend no FADD per element (measured: summing in the threads instead costs 4 % at 1024 keys).
  // d = 64: O only, the threads keep the row sums; that leaves tensor-memory room for 64-key tiles (2 x 64 S columns +
  // 2 x 64 O columns), half as many tiles and barrier round trips as the 32-key tiles the ones rows would force.
  static constexpr bool ONES = HD <= 32;
  static constexpr int NO = ONES ? HD + 16 : HD;
  static constexpr int Q_PLANE = kTaQ * QROW;
  static constexpr int K_PLANE = KT * QROW;
  static constexpr int VH_PLANE = NO * VROW;              // V^T hi rows (followed by the 16 rows of ones)
  static constexpr int VL_PLANE = HD * VROW;
  static constexpr int KV_STAGE = 2 * K_PLANE + VH_PLANE + VL_PLANE;
  static constexpr int KV_TX = 2 * K_PLANE + 2 * VL_PLANE;   // bytes one stage receives by TMA (the ones rows are constant)
  static constexpr int NKV = HD <= 32 ? 4 : 2;            // K/V ring depth
  // d <= 32: Q lives in TENSOR memory (the softmax threads copy their row there once; Q.K^T then takes its A operand from
  // TMEM like P.V does): an SS-mode M128 x N64 x K16 MMA is bound by its 6 KB of shared-memory operand reads (48 cycles
  // at 128 B/clk against 32 of tensor time, profiles/r02_ubench_mma_issue_rate.txt), and with two CTAs per SM the tensor
  // pipe -- not only the MUFU -- paces the tile loop.  d = 64 has no TMEM columns left for it and keeps Q in shared memory.
  static constexpr bool QTMEM = HD <= 32;
  static constexpr int Q_COLS = QTMEM ? HD : 0;            // hi words (HD / 2) | lo words (HD / 2)
  static constexpr int OFF_KV = QTMEM ? 0 : 2 * Q_PLANE;
  // half-merge exchange (128 rows x (HD + 2) floats) ALIASES the K/V ring: it is only touched after the last P.V MMA
  // has retired (every TMA load consumed, every MMA complete)
  static constexpr int OFF_X = OFF_KV;
  static_assert(kTaQ * (HD + 2) * 4 <= NKV * KV_STAGE, "exchange buffer must fit in the aliased region");
  static_assert(K_PLANE % (8 * QROW) == 0 && Q_PLANE % 1024 == 0 && VH_PLANE % 512 == 0 && KV_STAGE % 1024 == 0,
                "swizzle atom alignment");
  static constexpr int OFF_BAR = OFF_KV + NKV * KV_STAGE;
  static constexpr int N_BARS = 1 + 2 * NKV + 2 + 2 + 1 + 1;
  static constexpr size_t BYTES = 1024 + OFF_BAR + N_BARS * 8 + 16;
  static constexpr int TMEM_COLS = 256;                   // S / P double buffer (2 KT) + O per key half (2 NO)
  static_assert(2 * KT + 2 * NO + Q_COLS <= TMEM_COLS, "tensor memory budget (two CTAs per SM)");
  static_assert(2 * BYTES <= 232448, "two CTAs per SM");
};

struct TcAttnParams {
  float* msg;              // [rows][D] head-major message (pre-merge), fp32 -- or, when msg_hi is set,
  __half* msg_hi;          // fp16 hi / lo*2048 operand planes [rows][D] for the fused layer kernel (tc_gnn.cu)
  __half* msg_lo;
  int B, Np, D;
  const int* counts0; const int* counts1;
  int n_full0, n_full1;
  int cross;
  float scale_log2e;       // log2(e) / sqrt(d)
  int single;              // precision experiment (B200M_SINGLE=attn): hi planes only
  const __half* q_hi; const __half* q_lo;   // the q|k|v planes [rows][ldq] (Q read directly when it goes to tensor memory)
  int ldq; long long rows;
};

__device__ __forceinline__ unsigned long long add_f32x2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long splat_f32x2(float v) { return pack_f32x2(v, v); }

// EXPERIMENT, off in the product (B200M_ATTN_NPOLY = 0): measured on B200 at C2, moving 2 / 3 / 4 of a thread's 16
// pairs per tile from the MUFU to this polynomial made the kernel SLOWER (4.76 -> 5.09 / 5.23 / 5.28 ms per step): the
// tile loop is paced by instruction issue and dependency latency, not by MUFU throughput alone.
// (2^xa, 2^xb) on the FMA pipe instead of the MUFU: round-to-nearest range reduction x = n + f (magic-number add),
// degree-5 minimax polynomial for 2^f on [-0.5, 0.5] (max relative error 2.1e-7 evaluated in fp32 -- the class of
// ex2.approx's 2^-22), n added to the exponent field with integer arithmetic.  Packed fp32 instructions (two IEEE
// operations per issue slot).  Inputs are clamped to [-126, 127]: below, the result (~1e-38) packs to an fp16 zero exactly
// like ex2.approx.ftz's 0; above, 2^127 packs to an fp16 infinity, which raises the reference like an overflowed ex2.
#ifndef B200M_ATTN_NPOLY
#define B200M_ATTN_NPOLY 0            // pairs of a thread's 16 per key tile that take this path (0 = all MUFU)
#endif
__device__ __forceinline__ void exp2_poly_x2(unsigned long long x, float& ya, float& yb) {
  float xa, xb;
  unpack_f32x2(x, xa, xb);
  x = pack_f32x2(fminf(fmaxf(xa, -126.f), 127.f), fminf(fmaxf(xb, -126.f), 127.f));
  const unsigned long long t = add_f32x2(x, splat_f32x2(12582912.f));          // 1.5 * 2^23: integer part in the low mantissa bits
  const unsigned long long n = add_f32x2(t, splat_f32x2(-12582912.f));
  const unsigned long long f = fma_f32x2(n, splat_f32x2(-1.f), x);             // exact
  unsigned long long p = splat_f32x2(0.0013276472454890609f);
  p = fma_f32x2(p, f, splat_f32x2(0.009675540961325169f));
  p = fma_f32x2(p, f, splat_f32x2(0.05550713092088699f));
  p = fma_f32x2(p, f, splat_f32x2(0.24022120237350464f));
  p = fma_f32x2(p, f, splat_f32x2(0.6931469440460205f));
  p = fma_f32x2(p, f, splat_f32x2(1.0000001192092896f));
  float pa, pb, ta, tb;
  unpack_f32x2(p, pa, pb);
  unpack_f32x2(t, ta, tb);
  ya = __uint_as_float(__float_as_uint(pa) + (__float_as_uint(ta) << 23));
  yb = __uint_as_float(__float_as_uint(pb) + (__float_as_uint(tb) << 23));
}

// this thread's KH scores -> p = 2^(s c + neg) as packed fp16 hi / lo words; returns the packed maximum of the hi words
// SUM: also returns the sum of the p (otherwise the packed maximum of the hi words, as a float)
template <int KH, bool SUM>
__device__ __forceinline__ float softmax_words(const float* s, float c, float neg, uint32_t* hi, uint32_t* lo) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};                                    // four chains: the adds hide behind the MUFU
  __half2 pm = __float2half2_rn(0.f);
  const unsigned long long c2 = splat_f32x2(c), neg2 = splat_f32x2(neg);
#pragma unroll
  for (int i = 0; i < KH / 2; ++i) {
    float a, b;
    if (B200M_ATTN_NPOLY > 0 && (i * B200M_ATTN_NPOLY) % (KH / 2) < B200M_ATTN_NPOLY && KH == 32) {
      exp2_poly_x2(fma_f32x2(pack_f32x2(s[2 * i], s[2 * i + 1]), c2, neg2), a, b);
    } else {
      // (one packed fma.rn.f32x2 for the pair's two arguments measured slower as well: 4.45 -> 4.8 ms)
      a = fast_exp2(fmaf(s[2 * i], c, neg));                                // exp2(-inf) = 0 for masked keys
      b = fast_exp2(fmaf(s[2 * i + 1], c, neg));
    }
    if constexpr (SUM) acc[i & 3] += a + b;
    const uint32_t h = pack_f16x2(a, b);
    float ra, rb;
    residual_f16x2(h, a, b, ra, rb);
    hi[i] = h;
    lo[i] = pack_f16x2(ra, rb);
    if constexpr (!SUM) pm = __hmax2(pm, *reinterpret_cast<const __half2*>(&h));
  }
  if constexpr (SUM) return (acc[0] + acc[1]) + (acc[2] + acc[3]);
  else return __half2float(__hmax(__low2half(pm), __high2half(pm)));
}

template <int KH>
__device__ __forceinline__ float row_max(const float* s) {
  float m3[4] = {s[0], s[1], s[2], s[3]};                                 // 4 independent chains of 3-input maxima
#pragma unroll
  for (int i = 4; i + 1 < KH; i += 2) m3[(i >> 1) & 3] = fmaxf(m3[(i >> 1) & 3], fmaxf(s[i], s[i + 1]));
  return fmaxf(fmaxf(m3[0], m3[1]), fmaxf(m3[2], m3[3]));
}

#ifdef B200M_ATTN_TRACE
__device__ long long* g_attn_trace = nullptr;
#endif
template <int HD, int KT>
__global__ void __launch_bounds__(320, 2)
tc_attention_kernel(const __grid_constant__ CUtensorMap tm_q_hi, const __grid_constant__ CUtensorMap tm_q_lo,
                    const __grid_constant__ CUtensorMap tm_kv_hi, const __grid_constant__ CUtensorMap tm_kv_lo,
                    const __grid_constant__ CUtensorMap tm_vt_hi, const __grid_constant__ CUtensorMap tm_vt_lo,
                    TcAttnParams p) {
  using SM = TcAttnSmem<HD, KT>;
  constexpr int NO = SM::NO;
  constexpr int KH = KT / 2;          // keys (= S columns) per softmax thread and tile
  constexpr int PW = KH / 2;          // packed words per plane, thread and tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS / STS)
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + SM::OFF_KV;
  float* sX = reinterpret_cast<float*>(smem + SM::OFF_X);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* kv_full = q_full + 1;
  uint64_t* kv_empty = kv_full + SM::NKV;
  uint64_t* s_full = kv_empty + SM::NKV;     // [2] Q.K^T of the tile in buffer st retired
  uint64_t* p_full = s_full + 2;             // [2] the eight softmax warps wrote P into buffer st
  uint64_t* pv_done = p_full + 2;            // P.V of tile j retired (phase j); only waited on by a thread that is exactly
                                             // one phase behind (the rescale path) -- a parity wait cannot tell phase j from j-2
  uint64_t* o_done = pv_done + 1;            // the LAST P.V retired (single phase): the softmax warps run up to two tiles
                                             // ahead of the tensor pipe, so the final read must not use pv_done's parity
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // PDL: the key counts and the first Q / K / V^T loads of the set-up below already depend on earlier kernels of the
  // stream, so the wait comes first; what overlaps the predecessor's tail is this grid's launch and CTA scheduling
  pdl_trigger();
  pdl_wait();
#ifdef B200M_ATTN_TRACE
  long long* const tr = g_attn_trace ? g_attn_trace + 8 * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) : nullptr;
  if (tr && threadIdx.x == 64) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); tr[0] = sm; tr[1] = clock64(); }
#endif
  const int head = blockIdx.y;
  const int side = blockIdx.z / p.B, b = blockIdx.z - side * p.B;
  const int src = p.cross ? 1 - side : side;
  const int n_k = src == 0 ? (p.counts0 ? p.counts0[b] : p.n_full0) : (p.counts1 ? p.counts1[b] : p.n_full1);
  const int T = n_k > 0 ? cdiv(n_k, KT) : 0;
  const int q_row0 = (side * p.B + b) * p.Np + blockIdx.x * kTaQ;     // global row of the first query
  const int k_row0 = (src * p.B + b) * p.Np;
  const int cq = head * HD, ck = p.D + head * HD;          // first column of this head's q / k
  const int vt_row0 = (src * p.B + b) * p.D + head * HD;   // first row of this head in V^T [block][D][Np]

  // K / V^T tile j -> ring stage j % NKV (one elected thread)
  auto load_kv = [&](int j) {
    const int st = j % SM::NKV;
    mbar_expect_tx(&kv_full[st], SM::KV_TX);
    uint8_t* dst = sKV + st * SM::KV_STAGE;
    tma_load_2d(dst, &tm_kv_hi, &kv_full[st], ck, k_row0 + j * KT);
    tma_load_2d(dst + SM::K_PLANE, &tm_kv_lo, &kv_full[st], ck, k_row0 + j * KT);
    tma_load_2d(dst + 2 * SM::K_PLANE, &tm_vt_hi, &kv_full[st], j * KT, vt_row0);
    tma_load_2d(dst + 2 * SM::K_PLANE + SM::VH_PLANE, &tm_vt_lo, &kv_full[st], j * KT, vt_row0);
  };
  if (threadIdx.x == 0) {
    mbar_init(q_full, SM::QTMEM ? 8 : 1);       // Q landed: one TMA transaction, or the eight softmax warps' TMEM stores
    for (int i = 0; i < SM::NKV; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 8); }
    mbar_init(pv_done, 1);
    mbar_init(o_done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_q_hi); tma_prefetch_desc(&tm_q_lo);
    tma_prefetch_desc(&tm_kv_hi); tma_prefetch_desc(&tm_kv_lo);
    tma_prefetch_desc(&tm_vt_hi); tma_prefetch_desc(&tm_vt_lo);
    // the first loads go out BEFORE the tensor-memory allocation and the CTA-wide sync below: their L2 / DRAM latency
    // (~2.6 k cycles to the first score tile) overlaps the ~1.9 k cycles of set-up instead of following it
    // (Q and tile 0 do not wait for the key count either -- a pair without keys just lets them land, see below)
    if constexpr (!SM::QTMEM) {
      mbar_expect_tx(q_full, 2 * SM::Q_PLANE);
      tma_load_2d(sQ, &tm_q_hi, q_full, cq, q_row0);
      tma_load_2d(sQ + SM::Q_PLANE, &tm_q_lo, q_full, cq, q_row0);
    }
    load_kv(0);
  }
  if (warp == 1) tmem_alloc(tmem_slot, SM::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 2 * KT, tQ = tO + 2 * NO;
#ifdef B200M_ATTN_TRACE
  if (tr && threadIdx.x == 64) tr[2] = clock64();
#endif

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer (Q and tile 0 are on their way)
    for (int j = 1; j < T; ++j) {
      if (j >= SM::NKV) mbar_wait(&kv_empty[j % SM::NKV], ((j / SM::NKV) & 1) ^ 1);
      load_kv(j);
    }
    if (T == 0) {                      // nothing consumes the early loads: they must have landed before the CTA exits
      if constexpr (!SM::QTMEM) mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (converged warp)
    if (T > 0) {
      const uint32_t idesc_s = instr_desc(0, 128, KT);                    // fp16, A and B K-major
      const uint32_t idesc_o = instr_desc(0, 128, NO);                    // [O | l]
      const uint32_t idesc_v = instr_desc(0, 128, HD);                    // O only (P_hi x V_lo)
      const uint32_t q_base = smem_u32(sQ);
      auto issue_S = [&](int j) {
        const int st = j & 1;
        mbar_wait(&kv_full[j % SM::NKV], (j / SM::NKV) & 1);
        tc_fence_after();
        const uint32_t k_base = smem_u32(sKV + (j % SM::NKV) * SM::KV_STAGE);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks) {          // 16 channels = 32 B inside the swizzled row
            const uint64_t kh = smem_desc_sw<SM::QROW>(k_base + ks * 32);
            const uint64_t kl = smem_desc_sw<SM::QROW>(k_base + SM::K_PLANE + ks * 32);
            if constexpr (SM::QTMEM) {                  // 16 channels = 8 packed columns of the Q row in tensor memory
              const uint32_t qh = tQ + ks * 8, ql = tQ + HD / 2 + ks * 8;
              mma_f16_ts(tS + st * KT, qh, kh, idesc_s, ks != 0);
              if (!(kSingleExp && p.single)) {
                mma_f16_ts(tS + st * KT, qh, kl, idesc_s, 1);
                mma_f16_ts(tS + st * KT, ql, kh, idesc_s, 1);
              }
            } else {
              const uint64_t qh = smem_desc_sw<SM::QROW>(q_base + ks * 32);
              const uint64_t ql = smem_desc_sw<SM::QROW>(q_base + SM::Q_PLANE + ks * 32);
              mma_bf16(tS + st * KT, qh, kh, idesc_s, ks != 0);
              if (!(kSingleExp && p.single)) {
                mma_bf16(tS + st * KT, qh, kl, idesc_s, 1);
                mma_bf16(tS + st * KT, ql, kh, idesc_s, 1);
              }
            }
          }
          tc_commit(&s_full[st]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_S(0);
      if (T > 1) issue_S(1);
      for (int j = 0; j < T; ++j) {
        const int st = j & 1;
        mbar_wait(&p_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t v_base = smem_u32(sKV + (j % SM::NKV) * SM::KV_STAGE + 2 * SM::K_PLANE);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < KT / 16; ++ks) {           // 16 keys per MMA = 8 packed columns of P
            const int hf = ks / (KH / 16), kl = ks % (KH / 16);
            // key half hf owns S columns [hf KH, (hf + 1) KH): its P is [hi: PW columns | lo: PW columns] there
            const uint32_t a_hi = tS + st * KT + hf * KH + kl * 8, a_lo = a_hi + PW;
            const uint64_t vh = smem_desc_sw<SM::VROW>(v_base + ks * 32);
            const uint64_t vl = smem_desc_sw<SM::VROW>(v_base + SM::VH_PLANE + ks * 32);
            const uint32_t dO = tO + hf * NO;              // independent accumulators per key half
            const uint32_t acc = (j > 0) || kl != 0;       // accumulates across key tiles
            mma_f16_ts(dO, a_hi, vh, idesc_o, acc);
            if (!(kSingleExp && p.single)) {
              mma_f16_ts(dO, a_lo, vh, idesc_o, 1);
              mma_f16_ts(dO, a_hi, vl, idesc_v, 1);
            }
          }
          tc_commit(&kv_empty[j % SM::NKV]);
          tc_commit(pv_done);
          if (j == T - 1) tc_commit(o_done);
        }
        __syncwarp();
        // the next-but-one score tile overwrites P_j's columns: issued after P.V_j, executed after it (in-order pipe)
        if (j + 2 < T) issue_S(j + 2);
      }
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------------ softmax
    const int half = (warp - 2) >> 2;
    const int w4 = warp & 3;
    const int m = w4 * 32 + lane;
    const uint32_t lane_base = (uint32_t)(w4 * 32) << 16;
    const float c = p.scale_log2e;
    const uint32_t tS_mine = tS + lane_base + half * KH;
    const uint32_t tO_mine = tO + lane_base + half * NO;
    if constexpr (SM::QTMEM) {
      // this thread's query row -> tensor memory: half 0 copies the hi plane's HD / 2 packed words, half 1 the lo plane's
      // (memory order = the packed K-major A layout: word c of the row holds channels 2c, 2c + 1)
      if (T > 0) {
        constexpr int QW = HD / 2;
        uint32_t qw[QW];
        const long long qrow = (long long)q_row0 + m;
        if (qrow < p.rows) {
          const uint4* src = reinterpret_cast<const uint4*>((half ? p.q_lo : p.q_hi) + qrow * p.ldq + cq);
#pragma unroll
          for (int i = 0; i < QW / 4; ++i) {
            const uint4 v = __ldg(src + i);
            qw[4 * i] = v.x; qw[4 * i + 1] = v.y; qw[4 * i + 2] = v.z; qw[4 * i + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < QW; ++i) qw[i] = 0u;
        }
        tmem_st_words<QW>(tQ + lane_base + half * QW, qw);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(q_full);
      }
    }
    if constexpr (SM::ONES) {
      // the 16 constant rows of ones under every V^T hi tile (all elements equal, so the swizzle does not matter).  Written
      // here, while the first score tile is still on its way, instead of before the CTA-wide sync; the first P.V that
      // reads them is issued after this thread's p_full arrival.
      for (int i = threadIdx.x - 64; i < SM::NKV * 16 * SM::VROW / 16; i += 256) {
        const int stg = i / (16 * SM::VROW / 16), r = i - stg * (16 * SM::VROW / 16);
        *reinterpret_cast<uint4*>(sKV + stg * SM::KV_STAGE + 2 * SM::K_PLANE + HD * SM::VROW + r * 16) =
            make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
      }
      fence_proxy_async();
    }
    float m_ref = 0.f;                // reference the exponentials are taken against (log-2 domain after * c)
    float l_run = 0.f;                // sum of this key half's weights (every p is hi + lo to 2^-22: the sum the P.V MMAs apply)
    for (int j = 0; j < T; ++j) {
      const int st = j & 1;
      mbar_wait(&s_full[st], (j >> 1) & 1);
      tc_fence_after();
      float s[KH];
      tmem_ld_n<KH>(tS_mine + st * KT, s);
      const int kbase = j * KT + half * KH;
      if (kbase + KH > n_k) {
#pragma unroll
        for (int i = 0; i < KH; ++i)
          if (kbase + i >= n_k) s[i] = -INFINITY;
      }
#ifdef B200M_ATTN_TRACE
      if (tr && threadIdx.x == 64 && j == 0) tr[3] = clock64();
#endif
      if (j == 0) {                                   // the reference starts as the first tile's row maximum
        const float mx = row_max<KH>(s);
        m_ref = mx > -INFINITY ? mx : 0.f;
      }
      uint32_t hi[PW], lo[PW];
      // l_tile: the tile's row sum (d = 64) or its largest fp16 weight (d = 32, where the tensor core keeps the sums)
      float l_tile = softmax_words<KH, !SM::ONES>(s, c, -m_ref * c, hi, lo);
      // a weight above 2^8 (d = 64: a tile sum above 2^10, so no weight exceeds that; +inf if an exponential overflowed):
      // raise the reference and rescale what was accumulated so far
      const bool trig = l_tile > (SM::ONES ? 256.f : 1024.f);
      if (j > 0 && __any_sync(0xffffffffu, trig)) {
        // warp-uniform path (tcgen05.ld / .st are warp-collective); rows that did not trigger keep factor 1
        float f = 1.f;
        if (trig) {
          const float mx = row_max<KH>(s);
          f = fast_exp2((m_ref - mx) * c);
          m_ref = mx;
        }
        mbar_wait(pv_done, (j - 1) & 1);               // P.V of tile j-1 retired (it cannot be further: it needs our P_j)
        tc_fence_after();
        l_run *= f;
#pragma unroll 1
        for (int c0 = 0; c0 < NO; c0 += 16) {          // O of this key half, 16 columns at a time
          float ot[16];
          uint32_t ow[16];
          tmem_ld_n<16>(tO_mine + c0, ot);
#pragma unroll
          for (int i = 0; i < 16; ++i) ow[i] = __float_as_uint(ot[i] * f);
          tmem_st16(tO_mine + c0, ow);
          tmem_st_wait();      // tcgen05.st reads its source registers asynchronously: they are reused by the next chunk
        }
        l_tile = softmax_words<KH, !SM::ONES>(s, c, -m_ref * c, hi, lo);
      }
      if constexpr (!SM::ONES) l_run += l_tile;
      // P_j (A operand of the P.V MMA) over this thread's own S columns: [hi words | lo words]
      tmem_st_words<PW>(tS_mine + st * KT, hi);
      tmem_st_words<PW>(tS_mine + st * KT + PW, lo);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[st]);
    }
#ifdef B200M_ATTN_TRACE
    if (tr && threadIdx.x == 64) tr[4] = clock64();
#endif
    // ---- O of this key half: one TMEM read after the last P.V
    float o[NO];
#pragma unroll
    for (int i = 0; i < NO; ++i) o[i] = 0.f;
    if (T > 0) {
      mbar_wait(o_done, 0);
      tc_fence_after();
      tmem_ld_n<NO>(tO_mine, o);
      tc_fence_before();
    }
    if constexpr (SM::ONES) l_run = o[HD];            // sum of the weights, accumulated by the tensor core
    const float m_run = l_run > 0.f ? m_ref : -INFINITY;   // no valid key in this half: contributes nothing to the merge
    // ---- merge the two key halves
    float* xch = sX + (size_t)m * (HD + 2);
    if (half == 1) {
      xch[0] = m_run; xch[1] = l_run;
#pragma unroll
      for (int i = 0; i < HD; ++i) xch[2 + i] = o[i];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int row = blockIdx.x * kTaQ + m;
    if (half == 0 && row < p.Np) {
      const float mb = xch[0], lb = xch[1];
      const float mm = fmaxf(m_run, mb);
      const float ca = m_run == -INFINITY ? 0.f : fast_exp2((m_run - mm) * c);
      const float cb = mb == -INFINITY ? 0.f : fast_exp2((mb - mm) * c);
      const float l = l_run * ca + lb * cb;
      const float inv = l > 0.f ? 1.f / l : 0.f;
      const size_t off = ((size_t)(side * p.B + b) * p.Np + row) * p.D + head * HD;
      if (p.msg_hi) {
        // a thread owns a row: 32-byte stores (one full sector per instruction; the LSU's cost is per instruction and line)
        __half* dh = p.msg_hi + off;
        __half* dl = p.msg_lo + off;
        const bool wide32 = ((reinterpret_cast<uintptr_t>(dh) | reinterpret_cast<uintptr_t>(dl)) & 31) == 0;
#pragma unroll
        for (int g = 0; g < HD / 16; ++g) {
          uint32_t h8[8], l8[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float a = (o[16 * g + 2 * t] * ca + xch[2 + 16 * g + 2 * t] * cb) * inv;
            const float bb = (o[16 * g + 2 * t + 1] * ca + xch[2 + 16 * g + 2 * t + 1] * cb) * inv;
            h8[t] = pack_f16x2(a, bb);
            float ra, rb;
            residual_f16x2(h8[t], a, bb, ra, rb);
            l8[t] = pack_f16x2(ra * 2048.f, rb * 2048.f);
          }
          if (wide32) {
            st_global_256(dh + 16 * g, h8);
            st_global_256(dl + 16 * g, l8);
          } else {
            reinterpret_cast<uint4*>(dh + 16 * g)[0] = make_uint4(h8[0], h8[1], h8[2], h8[3]);
            reinterpret_cast<uint4*>(dh + 16 * g)[1] = make_uint4(h8[4], h8[5], h8[6], h8[7]);
            reinterpret_cast<uint4*>(dl + 16 * g)[0] = make_uint4(l8[0], l8[1], l8[2], l8[3]);
            reinterpret_cast<uint4*>(dl + 16 * g)[1] = make_uint4(l8[4], l8[5], l8[6], l8[7]);
          }
        }
      } else {
        float4* dst = reinterpret_cast<float4*>(p.msg + off);
#pragma unroll
        for (int g = 0; g < HD / 4; ++g) {
          float r4[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) r4[t] = (o[4 * g + t] * ca + xch[2 + 4 * g + t] * cb) * inv;
          dst[g] = make_float4(r4[0], r4[1], r4[2], r4[3]);
        }
      }
    }
  }
#ifdef B200M_ATTN_TRACE
  if (tr && threadIdx.x == 64) tr[5] = clock64();
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, SM::TMEM_COLS);
  }
#ifdef B200M_ATTN_TRACE
  if (tr && threadIdx.x == 64) tr[6] = clock64();
#endif
}

// 2-D view of a [rows][ld] fp16 plane; box = box_cols x box_rows with the swizzle that matches the row length
static bool make_f16_map(CUtensorMap* m, const void* base, size_t rows, size_t cols, size_t ld, int box_cols, int box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = box_cols * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : box_cols * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int HD, int KT>
static bool launch_tc_attn_t(LaunchCtx& ctx, const void* qkv_hi, const void* qkv_lo, const void* vt_hi,
                             const void* vt_lo, float* msg, int B, int Np,
                             int D, int heads, const int* c0, const int* c1, int nf0, int nf1, bool cross,
                             void* msg_hi, void* msg_lo, bool single) {
  ProfScope prof__(ctx, "tc_attention");
  const size_t rows = (size_t)2 * B * Np;
  CUtensorMap mq_hi, mq_lo, mk_hi, mk_lo, mv_hi, mv_lo;
  if (!make_f16_map(&mq_hi, qkv_hi, rows, 3 * D, 3 * D, HD, kTaQ) || !make_f16_map(&mq_lo, qkv_lo, rows, 3 * D, 3 * D, HD, kTaQ) ||
      !make_f16_map(&mk_hi, qkv_hi, rows, 3 * D, 3 * D, HD, KT) || !make_f16_map(&mk_lo, qkv_lo, rows, 3 * D, 3 * D, HD, KT) ||
      !make_f16_map(&mv_hi, vt_hi, (size_t)2 * B * D, Np, Np, KT, HD) || !make_f16_map(&mv_lo, vt_lo, (size_t)2 * B * D, Np, Np, KT, HD))
    return false;
  using SM = TcAttnSmem<HD, KT>;
  static SmemOptIn opt;
  auto kern = tc_attention_kernel<HD, KT>;
  if (!opt.ensure(kern, (int)SM::BYTES)) return false;
  TcAttnParams p;
  p.msg = msg; p.msg_hi = reinterpret_cast<__half*>(msg_hi); p.msg_lo = reinterpret_cast<__half*>(msg_lo); p.B = B; p.Np = Np; p.D = D; p.counts0 = c0; p.counts1 = c1; p.n_full0 = nf0; p.n_full1 = nf1;
  p.cross = cross ? 1 : 0;
  p.single = single ? 1 : 0;
  p.q_hi = reinterpret_cast<const __half*>(qkv_hi); p.q_lo = reinterpret_cast<const __half*>(qkv_lo);
  p.ldq = 3 * D; p.rows = (long long)rows;
  p.scale_log2e = 1.4426950408889634f / sqrtf((float)HD);
  dim3 grid(cdiv(Np, kTaQ), heads, 2 * B);
#ifdef B200M_ATTN_TRACE
  static long long* tbuf = nullptr;
  const int ncta = grid.x * grid.y * grid.z;
  static int n = 0;
  ++n;
  if (getenv("B200M_ATTN_TRACE") && n == 40) {
    cudaMalloc(&tbuf, (size_t)ncta * 64);
    cudaMemset(tbuf, 0, (size_t)ncta * 64);
    cudaMemcpyToSymbol(g_attn_trace, &tbuf, sizeof(tbuf));
  }
#endif
  launch_pdl(ctx, kPdlAttn, kern, grid, dim3(320), SM::BYTES, mq_hi, mq_lo, mk_hi, mk_lo, mv_hi, mv_lo, p);
  B200M_LAUNCH_CHECK(ctx, "tc_attention");
#ifdef B200M_ATTN_TRACE
  if (tbuf && n == 40) {
    std::vector<long long> hb((size_t)ncta * 8);
    cudaMemcpy(hb.data(), tbuf, hb.size() * 8, cudaMemcpyDeviceToHost);
    long long* nul = nullptr;
    cudaMemcpyToSymbol(g_attn_trace, &nul, sizeof(nul));
    // per SM: CTAs sorted by start; report mean phases and the idle gap between a CTA's end and the next start on the SM
    std::vector<std::vector<std::pair<long long, int>>> per_sm(256);
    for (int i = 0; i < ncta; ++i) per_sm[hb[8 * i] & 255].push_back({hb[8 * i + 1], i});
    double setup = 0, first = 0, loop = 0, epi = 0, exitc = 0, life = 0, span = 0; long long cnt = 0; int nsm = 0;
    for (auto& v : per_sm) {
      if (v.empty()) continue;
      ++nsm;
      std::sort(v.begin(), v.end());
      long long t_first = v.front().first, t_last = 0;
      for (auto& e : v) {
        const long long* r = &hb[8 * e.second];
        setup += r[2] - r[1]; first += r[3] - r[2]; loop += r[4] - r[3]; epi += r[5] - r[4]; exitc += r[6] - r[5]; life += r[6] - r[1];
        t_last = std::max(t_last, r[6]); ++cnt;
      }
      span += t_last - t_first;
    }
    fprintf(stderr, "ATTN trace: %d CTAs on %d SMs; per CTA mean cycles: setup %.0f, to first S %.0f, tile loop %.0f, output %.0f, exit sync %.0f, life %.0f; per SM span %.0f => %.0f per CTA slot (2 slots)\n",
            ncta, nsm, setup / cnt, first / cnt, loop / cnt, epi / cnt, exitc / cnt, life / cnt, span / nsm, span / nsm / (cnt / (double)nsm) * 2);
    // one SM in detail
    for (auto& v : per_sm) if (!v.empty()) { for (size_t i = 0; i < v.size() && i < 12; ++i) { const long long* r = &hb[8 * v[i].second]; fprintf(stderr, "  cta %5d start %8lld end %8lld\n", v[i].second, r[1] - v[0].first, r[6] - v[0].first); } break; }
  }
#endif
  return true;
}

// qkv_*: fp16 planes [2*B*Np][3D] (hi, lo = x - hi, unscaled); vt_*: fp16 planes [2*B][D][Np]
bool launch_tc_attention(LaunchCtx& ctx, const void* qkv_hi, const void* qkv_lo, const void* vt_hi,
                         const void* vt_lo, float* msg, int B, int Np, int D,
                         int heads, const int* counts0, const int* counts1, int n_full0, int n_full1, bool cross,
                         void* msg_hi, void* msg_lo, bool single) {
  const int hd = D / heads;
  if (Np % 8) return false;   // V^T rows must be 16-byte multiples for TMA
  if (hd == 16) return launch_tc_attn_t<16, 64>(ctx, qkv_hi, qkv_lo, vt_hi, vt_lo, msg, B, Np, D, heads, counts0, counts1, n_full0, n_full1, cross, msg_hi, msg_lo, single);
  if (hd == 32) return launch_tc_attn_t<32, 64>(ctx, qkv_hi, qkv_lo, vt_hi, vt_lo, msg, B, Np, D, heads, counts0, counts1, n_full0, n_full1, cross, msg_hi, msg_lo, single);
  if (hd == 64) return launch_tc_attn_t<64, 64>(ctx, qkv_hi, qkv_lo, vt_hi, vt_lo, msg, B, Np, D, heads, counts0, counts1, n_full0, n_full1, cross, msg_hi, msg_lo, single);
  return false;
}

}  // namespace b200m
