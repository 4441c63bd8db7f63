// One SuperGlue message-passing layer AFTER the attention, fused into a single tcgen05 kernel (D = 128):
//
//     msg   = W_merge att + b_merge                      (superglue_test.py:107, MultiHeadedAttention.merge)
//     hid   = relu(BN(W_1 [x ; msg] + b_1))              (:118-119, AttentionalPropagation.mlp, BatchNorm folded)
//     x    += W_2 hid + b_2                              (:119, :136 residual)
//     q|k|v = W_qkv' x + b_qkv'                          (:104-105 of the NEXT layer, heads de-interleaved)
//
// The unfused path runs these as four GEMM launches that round-trip msg, hid and x through HBM (~1 GB per layer at 64
// pairs) and re-split every activation tile per output-column tile.  Here a CTA owns 128 tokens and chains the four
// GEMMs on chip: activations stay in shared memory as fp16 hi/lo operand planes (fp16x3 split, see tc_gemm.cu /
// tc_conv.cu), accumulators in TMEM; HBM traffic per layer is the compulsory x read + x write + attention planes in
// + q|k|v planes out.
//
// Shared memory (1 CTA / SM):
//   XM   4 x 32 KB  activation K blocks (64 columns each, [hi plane | lo plane], 128 rows x 128 B, SWIZZLE_128B K-major):
//                   GEMM2 reads [x | msg], then hid overwrites all four, then x_new overwrites blocks 2,3;
//                   blocks 0,1 receive the NEXT tile's x (raw fp32 via TMA, split in place) while GEMM4 still runs.
//   RING 3 x 32 KB  operand tiles streamed by TMA: the attention planes (A of GEMM1) and every weight tile (B), the
//                   weights pre-tiled + pre-swizzled on the host in exactly the order they are consumed (1-D bulk copies).
// The V third of the q|k|v projection is issued with swapped operands (W_v tile as the M operand, x_new as the N operand),
// so it lands in TMEM transposed and the V^T planes the attention kernel streams are written with 16-byte stores.
// TMEM: two 256-column accumulator buffers (main | cross-term columns), alternated so an epilogue overlaps the next MMAs
//   where the data flow allows (GEMM4's three column tiles, GEMM1 of the next tile).
// Outputs never leave a thread as row-per-thread stores (32 different cache lines per warp instruction, measured: the
// LSU, not the tensor pipe, set the pace): x_new (fp32), the q / k planes and the V^T planes are staged in XM blocks
// 0,1 (free once GEMM3 retired) in the TMA box layout and written with cp.async.bulk.tensor stores; the next tile's x
// is fetched into the same blocks after the last store has read them.
// Warps: 0 = TMA producer, 1 = MMA issuer, 2..9 = epilogues / x splitter (thread = (token row, column half)).
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "kernels.cuh"
#include "tc_common.cuh"

namespace b200m {

using namespace tc;

constexpr int kGnPlane = 16384;                 // [128 rows][64 fp16]
constexpr int kGnTile = 2 * kGnPlane;           // hi + lo
constexpr int kGnRing = 3;
constexpr int kGnOffRing = 4 * kGnTile;
constexpr int kGnOffBar = kGnOffRing + kGnRing * kGnTile;
constexpr int kGnBars = 2 * kGnRing + 6 + 4 + 2;
constexpr int kGnOffBias = kGnOffBar + kGnBars * 8 + 16;
constexpr int kGnBiasFloats = 256 + 128 + 256;  // mlp1 (merge bias folded in) | mlp2 | next q,k  (the V bias is per TMEM lane)
constexpr size_t kGnSmem = kGnOffBias + kGnBiasFloats * 4;   // 232080 of the 232448 bytes a CTA may have: no alignment slack
static_assert(kGnSmem <= 232448, "shared memory budget");
constexpr float kGnLo = 2048.f;

__device__ __forceinline__ void gn_split8(const float* v, uint4& hi, uint4& lo, float lo_scale) {
  split8_f16(v, lo_scale, hi, lo);
}

// Developer aid (make EXTRA=-DB200M_GNN_TRACE, run with B200M_GNN_TRACE=1): clock64 stamps of one CTA's third tile for
// the producer / MMA / first epilogue warp, dumped to stderr by the launcher.  This is how the two stalls fixed in round 2
// were found (the residual's strided loads clogging the memory-instruction queue; the exposed x load at the tile start).
#ifdef B200M_GNN_TRACE
__device__ long long* g_gnn_trace = nullptr;
#define GT(slot) do { if (traced && lane == 0) tr[(slot)] = clock64(); } while (0)
#else
#define GT(slot) do { } while (0)
#endif
__global__ void __launch_bounds__(320, 1)
tc_gnn_layer_kernel(const __grid_constant__ CUtensorMap tm_att_hi, const __grid_constant__ CUtensorMap tm_att_lo,
                    const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_qkv_hi,
                    const __grid_constant__ CUtensorMap tm_qkv_lo, const __grid_constant__ CUtensorMap tm_vt_hi,
                    const __grid_constant__ CUtensorMap tm_vt_lo, GnnFusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];      // SWIZZLE_128B tiles need 1024-byte alignment (checked below)
  uint8_t* sXM = smem;
  uint8_t* sRing = smem + kGnOffRing;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kGnOffBar);
  uint64_t* full = bars;                    // ring tile landed
  uint64_t* empty = full + kGnRing;         // ring tile consumed by the MMAs
  uint64_t* x_full = empty + kGnRing;       // raw fp32 x tile landed in XM blocks 0,1
  uint64_t* x_free = x_full + 1;            // last staged store of the tile has read XM blocks 0,1: the next x may land
  uint64_t* x_ready = x_free + 1;           // x split into planes
  uint64_t* msg_ready = x_ready + 1;        // msg planes written (XM blocks 2,3)
  uint64_t* hid_ready = msg_ready + 1;      // hid planes written (XM blocks 0..3)
  uint64_t* xnew_ready = hid_ready + 1;     // x_new planes written (XM blocks 2,3)
  uint64_t* acc_full = xnew_ready + 1;      // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint64_t* xm01_free = acc_empty + 2;      // GEMM3's first two K blocks retired: XM blocks 0,1 may take the x reload
  uint64_t* x2_full = xm01_free + 1;        // the tile's fp32 x landed in XM blocks 0,1 a second time (residual add)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x2_full + 1);

  float* sBias = reinterpret_cast<float*>(smem + kGnOffBias);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef B200M_GNN_TRACE
  long long* const tr = g_gnn_trace;
  const bool traced_cta = tr != nullptr && blockIdx.x == 5;
  bool traced = false;
#define GT_ARM(cond) traced = traced_cta && (cond)
#else
#define GT_ARM(cond) do { } while (0)
#endif
  if (smem_u32(smem) & 1023) __trap();
  pdl_trigger();      // PDL: biases (weights), barriers and tensor memory are set up under the preceding kernel's tail
  for (int i = threadIdx.x; i < kGnBiasFloats; i += blockDim.x) sBias[i] = p.bias[i];
  const int ntiles = cdiv(p.rows, 128);
  const int n_wtiles = 12 + 2 * p.nt4;      // weight tiles per token tile after GEMM1: GEMM2 8, GEMM3 4, GEMM4 2 per column tile

  if (threadIdx.x == 0) {
    for (int i = 0; i < kGnRing; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(x_full, 1);
    mbar_init(x_free, 1);
    mbar_init(x_ready, 8);
    mbar_init(msg_ready, 8);
    mbar_init(hid_ready, 8);
    mbar_init(xnew_ready, 8);
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    mbar_init(xm01_free, 1);
    mbar_init(x2_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_att_hi); tma_prefetch_desc(&tm_att_lo); tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_qkv_hi); tma_prefetch_desc(&tm_qkv_lo); tma_prefetch_desc(&tm_vt_hi); tma_prefetch_desc(&tm_vt_lo);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();         // activations are read, and anything written, only from here on

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    int s = 0, ph = 0;
    auto load_x = [&](int tile, uint64_t* bar) {
      mbar_expect_tx(bar, 2 * kGnTile);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        tma_load_2d(sXM + kb * kGnTile, &tm_x, bar, kb * 64, tile * 128);
        tma_load_2d(sXM + kb * kGnTile + kGnPlane, &tm_x, bar, kb * 64 + 32, tile * 128);
      }
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      GT_ARM(it == 2);
      GT(0);
      if (it == 0) load_x(tile, x_full);
      const uint8_t* w = p.wts;
      for (int kb = 0; kb < 2; ++kb) {          // GEMM1: attention planes + W_merge tile per K block
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], kGnTile);
        tma_load_2d(sRing + s * kGnTile, &tm_att_hi, &full[s], kb * 64, tile * 128);
        tma_load_2d(sRing + s * kGnTile + kGnPlane, &tm_att_lo, &full[s], kb * 64, tile * 128);
        if (++s == kGnRing) { s = 0; ph ^= 1; }
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], kGnTile);
        bulk_load(sRing + s * kGnTile, w, kGnTile, &full[s]);
        w += kGnTile;
        if (++s == kGnRing) { s = 0; ph ^= 1; }
      }
      if (it > 0) {                             // this tile's x: XM blocks 0,1 are free once the previous tile's last
        GT(1);
        mbar_wait(x_free, (it - 1) & 1);        // staged store has read them (GEMM1 above does not need x)
        GT(2);
        load_x(tile, x_full);
      }
      for (int i = 0; i < n_wtiles; ++i) {      // GEMM2 (8), GEMM3 (4), GEMM4 (2 per column tile)
        GT(10 + 2 * i);
        mbar_wait(&empty[s], ph ^ 1);
        GT(11 + 2 * i);
        mbar_expect_tx(&full[s], kGnTile);
        bulk_load(sRing + s * kGnTile, w, kGnTile, &full[s]);
        w += kGnTile;
        if (++s == kGnRing) { s = 0; ph ^= 1; }
        if (i == 11) {
          // the residual needs the old x once more.  Fetching it with per-thread loads (one 1 KB-strided row per thread:
          // 4 k sector requests per tile) clogged the SM's memory-instruction queue exactly when the MMA warp issues
          // GEMM3 -- clock64 stamps showed its mbarrier / tcgen05.mma instructions taking 6 k cycles there.  TMA brings the
          // same 64 KB back into XM blocks 0,1 (free once GEMM3's K blocks 0,1 retired; the next weight tile waits for
          // the same event), in the box layout epilogue 3 updates in place.
          GT(3);
          mbar_wait(xm01_free, it & 1);
          GT(4);
          load_x(tile, x2_full);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, one lane issues)
    const uint32_t idesc = instr_desc(0 /*f16*/, 128, 128);
    const uint32_t idesc256 = instr_desc(0 /*f16*/, 128, 256);
    const uint32_t xm = smem_u32(sXM), ring = smem_u32(sRing);
    int s = 0, ph = 0;
    uint32_t use0 = 0, use1 = 0;              // scalars: a runtime-indexed array would live in local memory
    // one K block (64 columns): D_main += Ahi Whi ; D_cross += Ahi Wlo + Alo Whi.  The hi and lo planes of a weight
    // tile are contiguous (256 rows of 128 B), so [Whi ; Wlo] is one N = 256 operand whose product with Ahi lands as
    // [main | cross] in adjacent TMEM columns: two MMAs per K step instead of three (fewer shared-memory operand reads).
    // wide_n = false: `w` is the M-side operand's partner in the swapped (transposed) product and cannot be widened.
    auto block = [&](uint32_t d, uint32_t a, uint32_t w, bool first, bool wide_n) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t ah = smem_desc_sw128(a + ks * 32), al = smem_desc_sw128(a + kGnPlane + ks * 32);
        const uint64_t wh = smem_desc_sw128(w + ks * 32), wl = smem_desc_sw128(w + kGnPlane + ks * 32);
        const uint32_t acc = !(first && ks == 0);
        if (kSingleExp && p.single) {         // experiment: hi planes only
          mma_bf16(d, ah, wh, idesc, acc);
          continue;
        }
        if (wide_n) {
          mma_bf16(d, ah, wh, idesc256, acc);
        } else {
          mma_bf16(d, ah, wh, idesc, acc);
          mma_bf16(d + 128, ah, wl, idesc, acc);
        }
        mma_bf16(d + 128, al, wh, idesc, 1);
      }
    };
    auto acquire = [&](int b) {
      mbar_wait(&acc_empty[b], ((b ? use1 : use0) & 1) ^ 1);
      tc_fence_after();
    };
    // B tile from the ring, A from XM block `xb` (swapped: the ring tile is the M operand, XM the N operand, i.e. the
    // product comes out transposed -- used for V, whose consumer wants [channel][token])
    auto step_xm = [&](int b, int xb, bool first, bool swapped = false) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        // swapped: A = weight tile, B = x_new planes -- also hi plane | lo plane contiguous, so the wide form applies
        if (swapped) block(tmem_base + b * 256, ring + s * kGnTile, xm + xb * kGnTile, first, true);
        else block(tmem_base + b * 256, xm + xb * kGnTile, ring + s * kGnTile, first, true);
        tc_commit(&empty[s]);
      }
      __syncwarp();
      if (++s == kGnRing) { s = 0; ph ^= 1; }
    };
    auto publish = [&](int b) {
      if (elect_one()) tc_commit(&acc_full[b]);
      __syncwarp();
      if (b) ++use1; else ++use0;
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      GT_ARM(it == 2);
      GT(100);
      // ---- GEMM1 (merge) -> buffer 1
      acquire(1);
      GT(101);
      for (int kb = 0; kb < 2; ++kb) {
        const int sa = s, pa = ph;
        if (++s == kGnRing) { s = 0; ph ^= 1; }
        mbar_wait(&full[sa], pa);
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          block(tmem_base + 256, ring + sa * kGnTile, ring + s * kGnTile, kb == 0, true);
          tc_commit(&empty[sa]);
          tc_commit(&empty[s]);
        }
        __syncwarp();
        if (++s == kGnRing) { s = 0; ph ^= 1; }
      }
      publish(1);
      GT(102);
      // ---- GEMM2 (mlp layer 1), column tile 0 -> buffer 0: x part first, msg part once the merge epilogue is done
      mbar_wait(x_ready, it & 1);
      acquire(0);
      GT(103);
      step_xm(0, 0, true);
      step_xm(0, 1, false);
      GT(104);
      mbar_wait(msg_ready, it & 1);
      tc_fence_after();
      GT(105);
      step_xm(0, 2, false);
      step_xm(0, 3, false);
      publish(0);
      GT(106);
      // ---- GEMM2 column tile 1 -> buffer 1
      acquire(1);
      GT(107);
      for (int kb = 0; kb < 4; ++kb) step_xm(1, kb, kb == 0);
      publish(1);
      GT(108);
      // ---- GEMM3 (mlp layer 2) -> buffer 0
      mbar_wait(hid_ready, it & 1);
      acquire(0);
      GT(109);
      for (int kb = 0; kb < 4; ++kb) {
        step_xm(0, kb, kb == 0);
        if (kb == 1) {
          if (elect_one()) tc_commit(xm01_free);
          __syncwarp();
        }
      }
      publish(0);
      GT(110);
      // ---- GEMM4 (next layer's q|k|v), column tiles alternate buffers 0,1,0
      if (p.nt4 > 0) {
        mbar_wait(xnew_ready, it & 1);
        tc_fence_after();
        GT(111);
        for (int nt = 0; nt < p.nt4; ++nt) {
          const int b = nt & 1;
          acquire(b);
          step_xm(b, 2, true, nt == 2);
          step_xm(b, 3, false, nt == 2);
          publish(b);
          GT(112 + nt);
        }
      }
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------------ epilogues (thread = (row, column half))
    const int q4 = warp & 3;                       // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;
    const int row = q4 * 32 + lane;
    const int sw = row & 7;
    const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
    uint32_t use0 = 0, use1 = 0;
    const bool storer = threadIdx.x == 64;         // issues (and waits for) every staged TMA store of this CTA
    auto wait_acc = [&](int b) {
      mbar_wait(&acc_full[b], (b ? use1 : use0) & 1);
      tc_fence_after();
    };
    auto release_acc = [&](int b) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[b]);
      if (b) ++use1; else ++use0;
    };
    auto signal = [&](uint64_t* bar) {
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    auto epi_sync = [&](int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); };   // the 8 epilogue warps
    // staging protocol for XM blocks 0,1: stage_begin() before the first write of a stage (the previous stage's TMA
    // stores must have read the buffer), stage_end() after the last write (makes the writes visible to the async proxy
    // and lets the storer issue)
    auto stage_begin = [&]() {
      if (storer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      epi_sync(2);
    };
    auto stage_end = [&]() {
      fence_proxy_async();
      epi_sync(1);
    };
    // v[32] = main + cross / 2048 for columns [col0, col0 + 32) of buffer b
    auto load_acc = [&](int b, int col0, float* v) {
      float vc[32];
      const uint32_t t = tmem_base + lane_base + b * 256 + col0;
      tmem_ld32_issue(t, v);
      tmem_ld32_issue(t + 128, vc);
      tmem_ld_wait();
      if (!(kSingleExp && p.single)) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(vc[j], 1.f / kGnLo, v[j]);
      }
    };
    // 32 consecutive columns (starting at column c0 of the 64-column K block) of this row -> hi / lo planes
    auto store_planes = [&](uint8_t* kblock, int c0, const float* v, float lo_scale) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 hi, lo;
        gn_split8(v + 8 * g, hi, lo, lo_scale);
        const int off = row * 128 + ((((c0 >> 3) + g) ^ sw) << 4);
        *reinterpret_cast<uint4*>(kblock + off) = hi;
        *reinterpret_cast<uint4*>(kblock + kGnPlane + off) = lo;
      }
    };
    auto add_bias = [&](float* v, const float* bias) {      // 32 consecutive columns, 16-byte aligned, shared memory
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 bb = *(reinterpret_cast<const float4*>(bias) + g);
        v[4 * g] += bb.x; v[4 * g + 1] += bb.y; v[4 * g + 2] += bb.z; v[4 * g + 3] += bb.w;
      }
    };
    const float* b_mlp1 = sBias;             // W_1[:, 128:] b_merge + b_1 (the merge bias is folded on the host)
    const float* b_mlp2 = sBias + 256;
    const float* b_qk = sBias + 384;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      GT_ARM(it == 2 && warp == 2);
      GT(200);
      // ---- x: raw fp32 (two 32-column boxes per K block) -> hi / lo planes, in place; this thread: K block `half`
      mbar_wait(x_full, it & 1);
      GT(201);
      {
        uint8_t* kb = sXM + half * kGnTile;
        float e[64];
#pragma unroll
        for (int bx = 0; bx < 2; ++bx)
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(kb + bx * kGnPlane + row * 128 + ((q ^ sw) << 4));
            e[bx * 32 + q * 4 + 0] = v.x; e[bx * 32 + q * 4 + 1] = v.y;
            e[bx * 32 + q * 4 + 2] = v.z; e[bx * 32 + q * 4 + 3] = v.w;
          }
        __syncwarp();
        store_planes(kb, 0, e, kGnLo);
        store_planes(kb, 32, e + 32, kGnLo);
      }
      signal(x_ready);
      GT(202);
      // ---- epilogue 1: msg = acc (b_merge lives in b_1') -> XM blocks 2,3 (this thread: columns half*64 .. +63)
      wait_acc(1);
      if (it > 0 && p.nt4 > 0) stage_begin();     // the previous tile's V^T stores were staged in blocks 2,3: reads done?
      GT(203);
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        float v[32];
        load_acc(1, half * 64 + ch * 32, v);
        store_planes(sXM + (2 + half) * kGnTile, ch * 32, v, kGnLo);
      }
      release_acc(1);
      signal(msg_ready);
      GT(204);
      // ---- epilogue 2: hid = relu(acc + b_1); this thread: column tile `half` (buffer `half`), 128 columns.
      // hid overwrites [x | msg], so every GEMM2 MMA must have retired: wait for both buffers.
      wait_acc(0);
      wait_acc(1);
      GT(205);
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        float v[32];
        load_acc(half, ch * 32, v);
        add_bias(v, b_mlp1 + half * 128 + ch * 32);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        store_planes(sXM + (2 * half + (ch >> 1)) * kGnTile, (ch & 1) * 32, v, kGnLo);
      }
      release_acc(0);
      release_acc(1);
      signal(hid_ready);
      GT(206);
      // ---- epilogue 3: x_new = x + acc + b_2 -> operand planes in XM blocks 2,3 and, as fp32 in the TMA box layout of
      // the x load, into XM blocks 0,1 (free: GEMM3 has retired), from where one bulk tensor store writes the tile.
      // The old x arrives by TMA in exactly that box layout (second load of the tile, issued behind GEMM3) and is
      // updated in place.
      wait_acc(0);
      GT(207);
      mbar_wait(x2_full, it & 1);
      GT(216);
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        float v[32];
        load_acc(0, half * 64 + ch * 32, v);
        add_bias(v, b_mlp2 + half * 64 + ch * 32);
        uint8_t* box = sXM + half * kGnTile + ch * kGnPlane + row * 128;     // fp32 box: columns half*64 + ch*32 .. +31
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float4* cell = reinterpret_cast<float4*>(box + ((g ^ sw) << 4));
          const float4 old = *cell;
          v[4 * g] += old.x; v[4 * g + 1] += old.y; v[4 * g + 2] += old.z; v[4 * g + 3] += old.w;
          *cell = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        }
        store_planes(sXM + (2 + half) * kGnTile, ch * 32, v, kGnLo);
      }
      release_acc(0);
      signal(xnew_ready);
      GT(208);
      stage_end();
      if (storer) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          tma_store_2d(&tm_x, sXM + kb * kGnTile, kb * 64, tile * 128);
          tma_store_2d(&tm_x, sXM + kb * kGnTile + kGnPlane, kb * 64 + 32, tile * 128);
        }
        tma_store_commit();
      }
      __syncwarp();
      // ---- epilogue 4: next layer's q | k -> fp16 hi / lo planes [rows][384]; V arrives transposed (TMEM lane =
      // channel, column = token) -> V^T planes [block][128][Np].  Each 128-column tile is staged (block = 64 columns: hi
      // plane | lo plane, the layout of a SWIZZLE_128B box) and stored by TMA: q and k in XM blocks 0,1, the V^T tile in
      // blocks 2,3 (the x_new planes there are dead once GEMM4 retired) -- so blocks 0,1 are released for the NEXT tile's
      // x one column tile earlier and its DRAM latency hides behind the V epilogue (it was 4.6 k exposed cycles per tile).
      for (int nt = 0; nt < p.nt4; ++nt) {
        const int b = nt & 1;
        GT(209 + 2 * nt);
        wait_acc(b);
        GT(210 + 2 * nt);
        const float bv = nt == 2 ? __ldg(p.bias + kGnBiasFloats + row) : 0.f;   // V: this thread's channel = `row`
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          float v[32];
          load_acc(b, half * 64 + ch * 32, v);
          if (nt < 2) {
            add_bias(v, b_qk + nt * 128 + half * 64 + ch * 32);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += bv;
          }
          if (ch == 0) {
            stage_begin();                                        // every earlier staged store has read its buffer
            if (nt == 2 && storer) mbar_arrive(x_free);           // ... so blocks 0,1 may take the next tile's x now
          }
          store_planes(sXM + ((nt == 2 ? 2 : 0) + half) * kGnTile, ch * 32, v, 1.f);   // attention planes: unscaled residual
        }
        release_acc(b);
        stage_end();
        if (storer) {
          if (nt < 2) {
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
              tma_store_2d(&tm_qkv_hi, sXM + kb * kGnTile, nt * 128 + kb * 64, tile * 128);
              tma_store_2d(&tm_qkv_lo, sXM + kb * kGnTile + kGnPlane, nt * 128 + kb * 64, tile * 128);
            }
          } else {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {                       // 64-token runs (64 | Np: a run stays in one block)
              const int t0 = tile * 128 + hf * 64;
              if (t0 < p.rows) {
                const int blk = t0 / p.vt_np, rr0 = t0 - blk * p.vt_np;
                tma_store_2d(&tm_vt_hi, sXM + (2 + hf) * kGnTile, rr0, blk * 128);
                tma_store_2d(&tm_vt_lo, sXM + (2 + hf) * kGnTile + kGnPlane, rr0, blk * 128);
              }
            }
          }
          tma_store_commit();
        }
        __syncwarp();
      }
      GT(215);
      // ---- last layer (no q|k|v epilogue): blocks 0,1 may take the next tile's x once the x store has read them
      if (p.nt4 == 0 && storer) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(x_free);
      }
      __syncwarp();
    }
    if (storer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");    // all stores complete before exit
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static bool gn_map_f16(CUtensorMap* m, const void* base, size_t rows, int cols, int ld) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static bool gn_map_f32(CUtensorMap* m, const void* base, size_t rows, int cols, int ld) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, 128};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// att_hi / att_lo: attention output planes [rows][128] fp16 (lo scaled by 2048).  False if declined.
bool launch_tc_gnn_layer(LaunchCtx& ctx, const GnnFusedParams& p, const void* att_hi, const void* att_lo, int num_sms) {
  if (p.rows <= 0 || (p.nt4 != 0 && p.nt4 != 3) || p.ldx % 4) return false;
  if ((reinterpret_cast<uintptr_t>(p.X) | reinterpret_cast<uintptr_t>(p.wts) | reinterpret_cast<uintptr_t>(p.bias) |
       reinterpret_cast<uintptr_t>(p.qkv_hi) | reinterpret_cast<uintptr_t>(p.qkv_lo)) & 15)
    return false;
  ProfScope prof__(ctx, "tc_gnn_layer");
  if (p.vt_np % 64 || p.rows % 64) return false;
  CUtensorMap ma_hi, ma_lo, mx, mq_hi, mq_lo, mv_hi, mv_lo;
  const size_t vt_rows = (size_t)(p.rows / p.vt_np) * 128;
  if (!gn_map_f16(&ma_hi, att_hi, (size_t)p.rows, 128, 128) || !gn_map_f16(&ma_lo, att_lo, (size_t)p.rows, 128, 128) ||
      !gn_map_f32(&mx, p.X, (size_t)p.rows, 128, p.ldx) ||
      !gn_map_f16(&mq_hi, p.qkv_hi, (size_t)p.rows, 384, 384) || !gn_map_f16(&mq_lo, p.qkv_lo, (size_t)p.rows, 384, 384) ||
      !gn_map_f16(&mv_hi, p.vt_hi, vt_rows, p.vt_np, p.vt_np) || !gn_map_f16(&mv_lo, p.vt_lo, vt_rows, p.vt_np, p.vt_np))
    return false;
  static SmemOptIn opt;
  if (!opt.ensure(tc_gnn_layer_kernel, (int)kGnSmem)) return false;
  const int ntiles = cdiv(p.rows, 128);
  const int grid = ntiles < num_sms ? ntiles : num_sms;
#ifdef B200M_GNN_TRACE
  static long long* tbuf = nullptr;
  if (getenv("B200M_GNN_TRACE") && !tbuf) {
    cudaMalloc(&tbuf, 512 * 8);
    cudaMemset(tbuf, 0, 512 * 8);
    cudaMemcpyToSymbol(g_gnn_trace, &tbuf, sizeof(tbuf));
  }
#endif
  launch_pdl(ctx, kPdlGnn, tc_gnn_layer_kernel, dim3(grid), dim3(320), kGnSmem, ma_hi, ma_lo, mx, mq_hi, mq_lo, mv_hi, mv_lo, p);
  B200M_LAUNCH_CHECK(ctx, "tc_gnn_layer");
#ifdef B200M_GNN_TRACE
  if (tbuf) {
    static int n = 0;
    if (++n == 40) {
      long long hb[512];
      cudaMemcpy(hb, tbuf, sizeof(hb), cudaMemcpyDeviceToHost);
      const long long t0 = hb[200];
      fprintf(stderr, "GT producer: tile start %lld, x_free wait %lld -> %lld, x reload wait %lld -> %lld\n", hb[0] - t0,
              hb[1] - t0, hb[2] - t0, hb[3] - t0, hb[4] - t0);
      for (int i = 0; i < 18; ++i) fprintf(stderr, "GT producer wtile %2d: wait %6lld got %6lld\n", i, hb[10 + 2 * i] - t0, hb[11 + 2 * i] - t0);
      for (int i = 100; i <= 114; ++i) fprintf(stderr, "GT mma %d: %6lld\n", i, hb[i] - t0);
      for (int i = 200; i <= 216; ++i) fprintf(stderr, "GT epi %d: %6lld\n", i, hb[i] - t0);
    }
  }
#endif
  return true;
}

// ---------------------------------------------------------------- host-side weight stream
size_t gnn_fused_weight_floats(bool with_qkv) { return (size_t)(14 + (with_qkv ? 6 : 0)) * kGnTile / 4; }

// one 128-row x 64-column tile of a row-major [N][K] fp32 matrix -> hi plane | lo plane, SWIZZLE_128B K-major image
static void gn_pack_tile(const float* W, int K, int n0, int k0, uint8_t* dst) {
  __half* hi = reinterpret_cast<__half*>(dst);
  __half* lo = reinterpret_cast<__half*>(dst + kGnPlane);
  for (int n = 0; n < 128; ++n)
    for (int k = 0; k < 64; ++k) {
      const float w = W[(size_t)(n0 + n) * K + k0 + k];
      const __half h = __float2half_rn(w);
      const size_t idx = (size_t)n * 64 + ((((k >> 3) ^ (n & 7)) << 3) | (k & 7));
      hi[idx] = h;
      lo[idx] = __float2half_rn((w - __half2float(h)) * kGnLo);
    }
}

// W_merge [128][128], W_1 [256][256], W_2 [128][256], W_qkv [384][128] (or null): fp32 row-major, already permuted /
// BatchNorm-folded; tiles are emitted in the order the kernel consumes them.
void gnn_fused_pack_weights(const float* w_merge, const float* w1, const float* w2, const float* w_qkv, float* dst_f) {
  uint8_t* dst = reinterpret_cast<uint8_t*>(dst_f);
  for (int kb = 0; kb < 2; ++kb, dst += kGnTile) gn_pack_tile(w_merge, 128, 0, kb * 64, dst);
  for (int nt = 0; nt < 2; ++nt)
    for (int kb = 0; kb < 4; ++kb, dst += kGnTile) gn_pack_tile(w1, 256, nt * 128, kb * 64, dst);
  for (int kb = 0; kb < 4; ++kb, dst += kGnTile) gn_pack_tile(w2, 256, 0, kb * 64, dst);
  if (w_qkv)
    for (int nt = 0; nt < 3; ++nt)
      for (int kb = 0; kb < 2; ++kb, dst += kGnTile) gn_pack_tile(w_qkv, 128, nt * 128, kb * 64, dst);
}

}  // namespace b200m
