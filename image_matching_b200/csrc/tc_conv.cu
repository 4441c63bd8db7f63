// SuperPoint 3x3 convolutions on the 5th-generation tensor cores (tcgen05), fp32-class accuracy via a 2-term fp16
// split ("fp16x3"): v = hi + lo/2048 with hi = fp16(v), lo = fp16((v - hi) * 2048)  (representation error 2^-24 |v|,
// every fp16 x fp16 product is exact in the fp32 accumulator).  Compared with 3xTF32 (first version of this kernel)
// the operands are half as wide, so the shared-memory-bound MMA stream runs twice as fast, and the split is MORE
// accurate (tf32 hi/lo keeps 2^-22).  Range: activations must stay below the fp16 maximum (65504); the epilogue
// raises a sticky overflow flag that the host checks.
// Reference: superpoint/models/unet_parts.py:10-48, superpoint/models/superpoint_test.py:113-126.
//
// Implicit GEMM, one persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer   - per 32-channel K block ONE halo tile (18x18 pixels x 32 ch fp16, hi and lo planes) is
//                                fetched with cp.async.bulk.tensor (out-of-bounds = zero = the conv's zero padding);
//                                the nine taps are NOT re-fetched: a tap is just a +16 B / +288 B shift of the UMMA
//                                shared-memory descriptor's start address inside that tile (no-swizzle K-major
//                                canonical layout == the C4-planar activation layout).  Weight slabs [tap][kblock]
//                                arrive pre-laid-out with 1-D bulk copies.
//   warp 1      MMA issuer     - one thread issues tcgen05.mma.kind::f16, M=128 (16 rows x 8 px), N=NB, K=16;
//                                two accumulators (left / right 8-px half of the 16x16 tile) share every weight slab;
//                                D = Ahi*Bhi + (Ahi*Blo + Alo*Bhi)/2048  (fp32-class accuracy, keeps the detector
//                                logits stable enough for keypoint equality).  The tensor core truncates (RZ) when
//                                it adds a K=8 dot product into the fp32 accumulator, so the error grows with the
//                                number of accumulation steps: the small cross terms get their OWN accumulator
//                                (1/3 of the steps hit the large one; measured 3x lower error) and the two are
//                                summed with a rounded fp32 add in the epilogue.  NB=64: 2 x (2 px-halves x
//                                {main, cross} x 64) = 512 TMEM columns, double-buffered; NB=128: single-buffered.
//   warps 2..5  epilogue       - tcgen05.ld accumulators -> +bias, ReLU, optional 2x2 max-pool (warp shuffles),
//                                split into fp16 hi/lo planes for the next layer, coalesced 16-byte stores.
#include <cstring>
#include <cuda_fp16.h>
#include "kernels.cuh"
#include "tc_common.cuh"

namespace b200m {

using namespace tc;

constexpr int kTcTile = 16;                       // output tile is 16 x 16 pixels
constexpr int kTcKbGroups = 4;                    // 16-byte channel groups (8 fp16 channels) per K block -> 32 channels
constexpr int kTcUnitCh = 8;                      // channels per 16-byte unit
constexpr float kLoScale = 2048.f;                // lo planes / weights carry (v - hi) * 2^11

// KS = 3: 3x3 conv, 18x18 halo tile, nine taps by descriptor shifts.  KS = 1: the 1x1 heads (convPb / convDb),
// the same pipeline with a 16x16 tile and one tap per K block.
// RESIDENT: the layer's whole weight set (cin = 64, NB = 64, 3x3: 18 slabs = 144 KB) is loaded into shared memory once
// per persistent CTA instead of being re-streamed from L2 for every 16x16 tile.  Measured motivation: a 64 -> 64 layer
// needs 288 KB of weight slabs + 82 KB of activations per tile, i.e. ~34 B/clk/SM at tensor speed -- 80 % of the
// chip-wide L2 -> SM bandwidth -- and ran at ~50 % tensor-pipe utilisation.
template <int NB, int KS, bool FUSE1 = false, bool RESIDENT = false>
struct TcConvSmem {
  static constexpr int HALO = kTcTile + KS - 1;
  static constexpr int TAPS = KS * KS;
  static constexpr int PLANE_B = HALO * HALO * 16;        // bytes of one channel-group plane of the halo tile
  static constexpr int A_PLANE = kTcKbGroups * PLANE_B;   // hi (or lo) part of an A stage
  static constexpr int A_STAGE = 2 * A_PLANE;
  static constexpr int B_PLANE = kTcKbGroups * NB * 16;
  static constexpr int B_SLOT = 2 * B_PLANE;
  // weight-slab ring: a slab is consumed in 12 MMAs (~400-800 cycles) while an L2 fetch takes ~2-4k cycles, so the
  // ring must hold ~96 KB of slabs in flight (measured: 4 slots left the tensor pipe 60% idle waiting on B)
  static constexpr int NA = RESIDENT ? 2 : 3;          // A stages
  static constexpr int NBS = RESIDENT ? 2 * KS * KS : 98304 / B_SLOT;   // resident: every slab of the (cin = 64) layer
  static constexpr int BAR_OFF = NA * A_STAGE + NBS * B_SLOT;
  static constexpr int NBB = RESIDENT ? 0 : NBS;       // weight-ring barriers (none when the weights are resident)
  static constexpr int N_BARS = 2 * NA + 2 * NBB + 4 + 1;
  static constexpr int ACC_BUFS = NB == 64 ? 2 : 1;   // TMEM: bufs x 2 halves x {main, cross} x NB <= 512 columns
  // fused first layer: 20x20 image patch (fp32); the stem weights go from the kernel parameters into registers
  static constexpr int STEM_OFF = (BAR_OFF + N_BARS * 8 + 16 + 15) & ~15;
  static constexpr int STEM_BYTES = FUSE1 ? 400 * 4 : 0;
  static constexpr size_t BYTES = 128 /*align slack*/ + STEM_OFF + STEM_BYTES;
  static constexpr int STEM_WARPS = 8;
  static constexpr int THREADS = FUSE1 ? 192 + 32 * STEM_WARPS : 192;
};

// FUSE1: the layer's input is not read from memory but computed on the fly from the grayscale image: warps 6..13 run
// the network's first convolution (1 -> 64 channels, 3x3, BatchNorm folded, ReLU; unet_parts.py:10-24 first half) for
// the 18x18 halo of the tile on the CUDA cores and write it, already split into fp16 hi / lo planes, straight into the
// A stage the MMA warp consumes (same order of fp32 operations as conv1_direct -> bit-identical activations).  This
// removes the 157 MB / image-pair activation round trip through HBM of the unfused pair of kernels.
template <int NB, bool POOL, int KS, bool FUSE1, bool RESIDENT>
__global__ void __launch_bounds__(FUSE1 ? 448 : 192, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
               const __grid_constant__ TcConvParams p) {
  using SM = TcConvSmem<NB, KS, FUSE1, RESIDENT>;
  constexpr int kTcNB = SM::NBS, kTcNA = SM::NA;
  constexpr int kTcHalo = SM::HALO, kTcPlaneB = SM::PLANE_B, kTcAPlane = SM::A_PLANE, kTcAStage = SM::A_STAGE;
  constexpr int kTaps = SM::TAPS, kPad = KS / 2;
  extern __shared__ uint8_t smem_raw[];
  // (pointer arithmetic on the __shared__ array, not an integer round trip: keeps the address space, so every access
  // below compiles to LDS / STS instead of generic LD / ST)
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kTcNA * kTcAStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kTcNA;
  uint64_t* b_full = a_empty + kTcNA;
  uint64_t* b_empty = b_full + SM::NBB;
  uint64_t* acc_full = b_empty + SM::NBB;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_full = acc_empty + 2;          // RESIDENT: all weight slabs landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_x = cdiv(p.W, kTcTile), tiles_y = cdiv(p.H, kTcTile);
  const int ncb = p.cout_pad / NB;
  const int nkb = p.cin / (kTcKbGroups * kTcUnitCh);
  const int total = p.n * tiles_y * tiles_x * ncb;

  if (threadIdx.x == 0) {
    // a_full: one TMA transaction, or (fused stem) the four warps that fill the stage
    for (int i = 0; i < kTcNA; ++i) { mbar_init(&a_full[i], FUSE1 ? SM::STEM_WARPS / 2 : 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < SM::NBB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    mbar_init(w_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_hi);
    tma_prefetch_desc(&tm_lo);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above (barriers, tensor memory, descriptor prefetch) and the producer's weight requests below are
  // independent of the preceding kernel; activations are only read, and outputs only written, behind pdl_wait()
  pdl_trigger();
  if (!(warp == 0 && lane == 0)) pdl_wait();

  auto decode = [&](int tile, int& cb, int& x0, int& y0, int& img) {
    cb = tile % ncb;
    int t = tile / ncb;
    x0 = (t % tiles_x) * kTcTile;
    t /= tiles_x;
    y0 = (t % tiles_y) * kTcTile;
    img = t / tiles_y;
  };

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    int sa = 0, pa = 0, sb = 0, pb = 0;
    auto issue_A = [&](int tile, int kb) {
      int cb, x0, y0, img;
      decode(tile, cb, x0, y0, img);
      mbar_wait(&a_empty[sa], pa ^ 1);
      mbar_expect_tx(&a_full[sa], kTcAStage);
      uint8_t* dst = sA + sa * kTcAStage;
      tma_load_4d(dst, &tm_hi, &a_full[sa], kTcUnitCh * (x0 - kPad), y0 - kPad, p.in_c8_off + kb * kTcKbGroups, img);
      tma_load_4d(dst + kTcAPlane, &tm_lo, &a_full[sa], kTcUnitCh * (x0 - kPad), y0 - kPad,
                  p.in_c8_off + kb * kTcKbGroups, img);
      if (++sa == kTcNA) { sa = 0; pa ^= 1; }
    };
    if (RESIDENT) {            // the whole layer (ncb == 1, nkb * taps == NBS slabs), once
      mbar_expect_tx(w_full, kTcNB * SM::B_SLOT);
      for (int i = 0; i < kTcNB; ++i)
        bulk_load(sB + i * SM::B_SLOT, reinterpret_cast<const uint8_t*>(p.wpk) + (size_t)i * SM::B_SLOT, SM::B_SLOT, w_full);
    }
    pdl_wait();
    if (!FUSE1 && (int)blockIdx.x < total) issue_A(blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const int cb = tile % ncb;
      for (int kb = 0; kb < nkb; ++kb) {
        // keep the activation halo one K block ahead of the weight slabs
        if (!FUSE1) {
          if (kb + 1 < nkb) issue_A(tile, kb + 1);
          else if (tile + (int)gridDim.x < total) issue_A(tile + gridDim.x, 0);
        }
        if (RESIDENT) continue;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpk) + (size_t)(cb * nkb + kb) * kTaps * SM::B_SLOT;
        for (int tap = 0; tap < kTaps; ++tap) {
          mbar_wait(&b_empty[sb], pb ^ 1);
          mbar_expect_tx(&b_full[sb], SM::B_SLOT);
          bulk_load(sB + sb * SM::B_SLOT, wsrc + (size_t)tap * SM::B_SLOT, SM::B_SLOT, &b_full[sb]);
          if (++sb == kTcNB) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, one elected lane issues)
    // Two MMAs per product instead of three: the weight slab keeps, per 16-byte K chunk, the NB hi rows followed by the
    // NB lo rows, so [W_hi ; W_lo] is ONE N = 2*NB operand and  A_hi x [W_hi ; W_lo]^T  lands as [main | cross] in
    // adjacent TMEM columns; only A_lo x W_hi^T remains.  Same tensor-pipe time, but 14 (NB = 64) / 20 (NB = 128) KB of
    // shared-memory operand reads per K step instead of 18 / 24 -- the N = 64 / 128 MMAs were shared-memory-bandwidth
    // bound (A 4 KB + B 2..4 KB per 34..68 tensor cycles at 128 B/clk).
    const uint32_t idesc2 = instr_desc(0 /*f16*/, 128, 2 * NB);
    const uint32_t idesc1 = instr_desc(0 /*f16*/, 128, NB);
    const uint64_t a_hi32 = (smem_desc_nosw(0, kTcPlaneB, kTcHalo * 16) >> 32) << 32;
    const uint32_t a_lo16 = (uint32_t)((kTcPlaneB >> 4) << 16);
    const uint64_t b_hi32 = (smem_desc_nosw(0, 2 * NB * 16, 128) >> 32) << 32;
    const uint32_t b_lo16 = (uint32_t)(((2 * NB * 16) >> 4) << 16);
    int sa = 0, pa = 0, sb = 0, pb = 0, lt = 0;
    if (RESIDENT) mbar_wait(w_full, 0);
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++lt) {
      const int buf = SM::ACC_BUFS == 2 ? (lt & 1) : 0;
      const int aph = SM::ACC_BUFS == 2 ? ((lt >> 1) & 1) : (lt & 1);
      mbar_wait(&acc_empty[buf], aph ^ 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + buf * (4 * NB);
      const bool single = kSingleExp && (tile % ncb) >= p.single_from_cb;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&a_full[sa], pa);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + sa * kTcAStage);
        for (int tap = 0; tap < kTaps; ++tap) {
          if (!RESIDENT) {
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
          }
          const uint32_t b_base = smem_u32(sB + (RESIDENT ? kb * kTaps + tap : sb) * SM::B_SLOT);
          const int ky = tap / KS, kx = tap - KS * ky;
          if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t b_off = b_base + ks * 2 * (2 * NB * 16);     // K chunks 2ks, 2ks+1: [hi rows | lo rows] each
            const uint64_t bw = b_hi32 | (uint64_t)(b_lo16 | ((b_off >> 4) & 0x3FFF));
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              const uint32_t a_off = a_base + ks * 2 * kTcPlaneB + (ky * kTcHalo + kx + sub * 8) * 16;
              const uint64_t ah = a_hi32 | (uint64_t)(a_lo16 | ((a_off >> 4) & 0x3FFF));
              const uint64_t al = a_hi32 | (uint64_t)(a_lo16 | (((a_off + kTcAPlane) >> 4) & 0x3FFF));
              const uint32_t d = d0 + sub * (2 * NB);      // main accumulator; cross accumulator at d + NB
              const uint32_t acc = (kb | tap | ks) != 0;
              if (single) {
                mma_bf16(d, ah, bw, idesc1, acc);       // experiment: A_hi W_hi^T only
              } else {
                mma_bf16(d, ah, bw, idesc2, acc);       // kind::f16: [A_hi W_hi^T | A_hi W_lo^T]
                mma_bf16(d + NB, al, bw, idesc1, 1);    // + A_lo W_hi^T (first NB rows of the chunk)
              }
            }
          }
          if (!RESIDENT) tc_commit(&b_empty[sb]);
          if (tap == kTaps - 1) tc_commit(&a_empty[sa]);
          if (tap == kTaps - 1 && kb == nkb - 1) tc_commit(&acc_full[buf]);
          }
          __syncwarp();
          if (++sb == kTcNB) { sb = 0; pb ^= 1; }
        }
        if (++sa == kTcNA) { sa = 0; pa ^= 1; }
      }
    }
  } else if (FUSE1 && warp >= 6) {
    // ------------------------------------------------------------------ stem: first convolution -> A stages
    float* patch = reinterpret_cast<float*>(smem + SM::STEM_OFF);       // 20 x 20 image pixels around the tile
    // A warp owns ONE 8-channel unit (8 warps = the 64 channels = both 32-channel A stages of a tile) and keeps that
    // unit's 72 weights + 8 biases in registers (read once from the kernel parameters); lane = halo pixel.  Warps 0-3
    // fill the tile's first A stage, warps 4-7 the second one, concurrently.
    constexpr int kStemT = 32 * SM::STEM_WARPS;
    static_assert(SM::STEM_WARPS == 8, "one stem warp per 8-channel unit");
    const int t = threadIdx.x - 192;                                     // 0..kStemT-1
    const int unit = t >> 5;                                             // 8-channel unit 0..7
    const int kbw = unit >> 2, g = unit & 3;                             // A stage of the tile / unit inside the stage
    unsigned long long wr2[9][4];      // (weight of channel 2c, weight of channel 2c + 1) pairs
    float br[8];
#pragma unroll
    for (int tp = 0; tp < 9; ++tp)
#pragma unroll
      for (int c2 = 0; c2 < 4; ++c2)
        wr2[tp][c2] = pack_f32x2(p.c1[tp * 64 + unit * 8 + 2 * c2], p.c1[tp * 64 + unit * 8 + 2 * c2 + 1]);
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) br[ch] = p.c1[9 * 64 + unit * 8 + ch];
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      int cb, x0, y0, img;
      decode(tile, cb, x0, y0, img);
      asm volatile("bar.sync 3, %0;" ::"n"(kStemT) : "memory");          // everyone is done with the previous patch
      const size_t img_off = (size_t)img * p.H * p.W;
      for (int i = t; i < 400; i += kStemT) {
        const int r = i / 20, c = i - r * 20;
        const int gy = y0 - 2 + r, gx = x0 - 2 + c;
        float v = 0.f;
        if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
          const size_t o = img_off + (size_t)gy * p.W + gx;
          // 8-bit input: the loader's `pixel / 255.` (float64) followed by .float() == one correctly rounded fp32 division
          v = p.img_u8 ? __fdiv_rn((float)__ldg(p.img_u8 + o), 255.f) : __ldg(p.img + o);
        }
        patch[i] = v;
      }
      asm volatile("bar.sync 3, %0;" ::"n"(kStemT) : "memory");
      const int cst = 2 * it + kbw;                                      // running A-stage index (two per tile)
      const int sa = cst % kTcNA, pa = (cst / kTcNA) & 1;
      mbar_wait(&a_empty[sa], pa ^ 1);
      uint8_t* dst = sA + sa * kTcAStage + g * kTcPlaneB;
#pragma unroll 1
      for (int px = lane; px < kTcHalo * kTcHalo; px += 32) {
        const int py = px / kTcHalo, pxx = px - py * kTcHalo;
        const int gy = y0 - 1 + py, gx = x0 - 1 + pxx;
        uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;                  // outside the image: the NEXT conv's zero padding
        if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
          // packed fp32 FMAs (fma.rn.f32x2: two IEEE fp32 FMAs per instruction, same results): 36 instead of 72 issue
          // slots per pixel and unit -- the stem warps share their schedulers with the MMA-issuing and epilogue warps
          unsigned long long e2[4];
#pragma unroll
          for (int c2 = 0; c2 < 4; ++c2) e2[c2] = pack_f32x2(br[2 * c2], br[2 * c2 + 1]);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const float v = patch[(py + ky) * 20 + pxx + kx];
              const unsigned long long vv = pack_f32x2(v, v);
#pragma unroll
              for (int c2 = 0; c2 < 4; ++c2) e2[c2] = fma_f32x2(vv, wr2[ky * 3 + kx][c2], e2[c2]);
            }
          float e[8];
#pragma unroll
          for (int c2 = 0; c2 < 4; ++c2) unpack_f32x2(e2[c2], e[2 * c2], e[2 * c2 + 1]);
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) e[ch] = fmaxf(e[ch], 0.f);
          split8_f16(e, kLoScale, hi, lo);
        }
        *reinterpret_cast<uint4*>(dst + px * 16) = hi;
        *reinterpret_cast<uint4*>(dst + kTcAPlane + px * 16) = lo;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[sa]);
    }
  } else if (warp >= 2 && warp < 6) {
    // ------------------------------------------------------------------ epilogue (128 threads = 128 TMEM lanes)
    const int w4 = warp & 3;
    const int m = w4 * 32 + lane;
    const int xs = m & 7, ys = m >> 3;
    const int Ho = POOL ? p.H / 2 : p.H, Wo = POOL ? p.W / 2 : p.W;
    const size_t oplane = (size_t)Ho * Wo;
    int lt = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++lt) {
      int cb, x0, y0, img;
      decode(tile, cb, x0, y0, img);
      const int buf = SM::ACC_BUFS == 2 ? (lt & 1) : 0;
      const int aph = SM::ACC_BUFS == 2 ? ((lt >> 1) & 1) : (lt & 1);
      mbar_wait(&acc_full[buf], aph);
      tc_fence_after();
      const bool single = kSingleExp && cb >= p.single_from_cb;                 // experiment: no cross accumulator
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {
        const int x = x0 + sub * 8 + xs, y = y0 + ys;
        int X, Y;
        bool writer;
        if (POOL) {
          X = x >> 1; Y = y >> 1;
          writer = !(xs & 1) && !(ys & 1) && (Y < Ho) && (X < Wo);
        } else {
          X = x; Y = y;
          writer = (y < p.H) && (x < p.W);
        }
        float ss = 0.f;                       // sum of squares over this tile's channels (descriptor head only)
#pragma unroll 1
        for (int ch = 0; ch < NB / 32; ++ch) {
          float v[32], vc[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(w4 * 32) << 16) + buf * (4 * NB) + sub * (2 * NB) + ch * 32;
          tmem_ld32(taddr, v);
          tmem_ld32(taddr + NB, vc);
          const int c0 = cb * NB + ch * 32;
          // bias: from the kernel parameters (constant bank, one uniform read) when the layer has <= 128 output channels,
          // else 32 L1 loads per chunk
          float bj[32];
          if (p.bias_in_params) {
#pragma unroll
            for (int j = 0; j < 32; ++j) bj[j] = p.bias_c[c0 + j];
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) bj[j] = __ldg(p.bias + c0 + j);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float t = (single ? v[j] : fmaf(vc[j], 1.f / kLoScale, v[j])) + bj[j];
            if (p.relu) t = fmaxf(t, 0.f);
            if (POOL) {
              t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, 1));
              t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, 8));
            }
            v[j] = t;
          }
          if (writer) {
            if (p.out_lo) {
              // fp16 hi / lo planes, C8-planar: 4 units of 8 channels per 32-column chunk
              const size_t g0 = (size_t)img * p.out_c4_total + p.out_c4_off + (c0 >> 3);
              uint4* oh = reinterpret_cast<uint4*>(p.out_hi) + g0 * oplane + (size_t)Y * Wo + X;
              uint4* ol = reinterpret_cast<uint4*>(p.out_lo) + g0 * oplane + (size_t)Y * Wo + X;
              bool ovf = false;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float mx = 0.f;
#pragma unroll
                for (int j = 0; j < 8; j += 2) mx = fmaxf(mx, fmaxf(fabsf(v[8 * g + j]), fabsf(v[8 * g + j + 1])));
                ovf |= mx > 65000.f;
                uint4 h4, l4;
                split8_f16(v + 8 * g, kLoScale, h4, l4);
                oh[(size_t)g * oplane] = h4;
                ol[(size_t)g * oplane] = l4;
              }
              if (ovf && p.overflow) *p.overflow = 1;
            } else {
              // full fp32, C4-planar (detector logits / raw descriptors)
              if (p.sumsq) {
#pragma unroll
                for (int j = 0; j < 32; ++j) ss = fmaf(v[j], v[j], ss);      // channel order: same chain as a serial pass
              }
              const size_t g0 = (size_t)img * p.out_c4_total + p.out_c4_off + (c0 >> 2);
              float4* oh = reinterpret_cast<float4*>(p.out_hi) + g0 * oplane + (size_t)Y * Wo + X;
#pragma unroll
              for (int g = 0; g < 8; ++g)
                oh[(size_t)g * oplane] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
            }
          }
        }
        if (p.sumsq && writer && !p.out_lo)
          p.sumsq[((size_t)(img * ncb + cb) * p.H + Y) * p.W + X] = ss;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static bool make_act_map(CUtensorMap* m, const void* base, int n, int c4, int H, int W, int halo) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)kTcUnitCh * W, (cuuint64_t)H, (cuuint64_t)c4, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)c4 * H * W * 16};
  cuuint32_t box[4] = {(cuuint32_t)(kTcUnitCh * halo), (cuuint32_t)halo, kTcKbGroups, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int NB, bool POOL, int KS, bool FUSE1 = false, bool RESIDENT = false>
static bool launch_tc_t(LaunchCtx& ctx, const TcConvParams& p, int num_sms) {
  using SM = TcConvSmem<NB, KS, FUSE1, RESIDENT>;
  static_assert(SM::BYTES <= 232448, "shared memory budget");
  ProfScope prof__(ctx, KS == 3 ? (FUSE1 ? "tc_conv3x3_stem" : "tc_conv3x3") : "tc_conv1x1");
  const int c8_total = p.in_c8_total > 0 ? p.in_c8_total : p.cin / kTcUnitCh;
  CUtensorMap tm_hi, tm_lo;
  if (FUSE1) {
    memset(&tm_hi, 0, sizeof(tm_hi));   // the activation maps are not used: the stem warps produce the A operand
    memset(&tm_lo, 0, sizeof(tm_lo));
  } else {
    if (!make_act_map(&tm_hi, p.in_hi, p.n, c8_total, p.H, p.W, SM::HALO)) return false;
    if (!make_act_map(&tm_lo, p.in_lo, p.n, c8_total, p.H, p.W, SM::HALO)) return false;
  }
  static SmemOptIn opt;
  auto kern = tc_conv_kernel<NB, POOL, KS, FUSE1, RESIDENT>;
  if (!opt.ensure(kern, (int)SM::BYTES)) return false;
  const int total = p.n * cdiv(p.H, kTcTile) * cdiv(p.W, kTcTile) * (p.cout_pad / NB);
  const int grid = total < num_sms ? total : num_sms;
  launch_pdl(ctx, kPdlConv, kern, dim3(grid), dim3(SM::THREADS), SM::BYTES, tm_hi, tm_lo, p);
  B200M_LAUNCH_CHECK(ctx, KS == 3 ? "tc_conv3x3" : "tc_conv1x1");
  return true;
}

bool launch_tc_conv(LaunchCtx& ctx, const TcConvParams& p, int num_sms) {
  if (p.cin % 32 || p.cout_pad % p.nb || (p.nb != 64 && p.nb != 128) || (p.ks != 1 && p.ks != 3)) return false;
  if (p.ks == 1) {
    if (p.pool) return false;
    return p.nb == 64 ? launch_tc_t<64, false, 1>(ctx, p, num_sms) : launch_tc_t<128, false, 1>(ctx, p, num_sms);
  }
  if (p.img || p.img_u8) {   // fused first layer (image -> 64 channels) in front of a 64 -> 64 pooled layer
    if (p.nb != 64 || p.cin != 64 || p.cout_pad != 64 || !p.pool) return false;
    // (weights resident, 2 A stages = one tile: the two stages are filled concurrently by stem warps 0-3 / 4-7;
    // measured 7.6 ms vs 8.7 ms with streamed weights + 3 stages)
    return launch_tc_t<64, true, 3, true, true>(ctx, p, num_sms);
  }
  if (p.nb == 64 && p.cin == 64 && p.cout_pad == 64)     // 64 -> 64 layers: weights resident in shared memory
    return p.pool ? launch_tc_t<64, true, 3, false, true>(ctx, p, num_sms) : launch_tc_t<64, false, 3, false, true>(ctx, p, num_sms);
  if (p.nb == 64) return p.pool ? launch_tc_t<64, true, 3>(ctx, p, num_sms) : launch_tc_t<64, false, 3>(ctx, p, num_sms);
  return p.pool ? launch_tc_t<128, true, 3>(ctx, p, num_sms) : launch_tc_t<128, false, 3>(ctx, p, num_sms);
}

// size of the packed weights in floats (the arena is float-typed; the content is fp16)
size_t tc_conv_weight_floats(int cin, int cout_pad, int nb, int ks) {
  return (size_t)(cout_pad / nb) * (cin / 32) * ks * ks * 2 * kTcKbGroups * nb * 4;
}

// Host-side weight packing: w[cout][cin][ks][ks] (BatchNorm already folded) ->
// [cout_blk][kblock(32 ch)][tap][unit of 8 ch][plane hi/lo][n][8 halves], i.e. the exact shared-memory image of a
// B slab (per K chunk the nb hi rows then the nb lo rows: one N = 2*nb operand); hi = fp16(w), lo = fp16((w - hi) * 2048).
void tc_conv_pack_weights(const double* w, int cout, int cin, int cout_pad, int nb, int ks, float* dst_f) {
  __half* dst = reinterpret_cast<__half*>(dst_f);
  const int ncb = cout_pad / nb, nkb = cin / 32, taps = ks * ks;
  for (int cb = 0; cb < ncb; ++cb)
    for (int kb = 0; kb < nkb; ++kb)
      for (int tap = 0; tap < taps; ++tap)
        for (int kc = 0; kc < kTcKbGroups; ++kc)
          for (int n = 0; n < nb; ++n)
            for (int j = 0; j < kTcUnitCh; ++j) {
              const int o = cb * nb + n, ci = kb * 32 + kc * kTcUnitCh + j;
              const float v = o < cout ? (float)w[((size_t)o * cin + ci) * taps + tap] : 0.f;
              const __half hi = __float2half_rn(v);
              const __half lo = __float2half_rn((v - __half2float(hi)) * kLoScale);
              const size_t slab = ((size_t)(cb * nkb + kb) * taps + tap) * 2 * kTcKbGroups * nb * kTcUnitCh;
              dst[slab + (((size_t)kc * 2 + 0) * nb + n) * kTcUnitCh + j] = hi;
              dst[slab + (((size_t)kc * 2 + 1) * nb + n) * kTcUnitCh + j] = lo;
            }
}

}  // namespace b200m
