// Internal launcher interface between api.cu (orchestration) and the kernel translation units.
// Activation layout used by every SuperPoint kernel ("C4-planar"): [image][C/4][H][W][4] fp32,
// i.e. pixels of a 4-channel group are contiguous float4s.  It makes (a) halo tiles of a channel
// group a dense 2-D box (one TMA box / coalesced float4 rows), (b) a shifted 3x3 tap a plain
// +16 B / +pitch address offset, and (c) the epilogue store of 4 consecutive output channels of
// 32 consecutive pixels one 512 B coalesced row.
// SuperGlue token layout: [side][pair][token][ld] fp32 row-major ("token-major"), K contiguous.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace b200m {

// ------------------------------------------------------------------ SuperPoint dense (sp_conv.cu)
struct ConvParams {
  const float* in;     // C4-planar, in_c4_total groups per image
  int in_c4_total;
  int in_c4_off;       // first input channel group read
  int cin;             // input channels (multiple of 8)
  const float* wpk;    // packed [cout_blk][cin_chunk][tap][8][64]
  const float* bias;   // [cout_pad]
  float* out;          // C4-planar, out_c4_total groups per image
  int out_c4_total;
  int out_c4_off;
  int cout_pad;        // multiple of 64
  int n, H, W;         // input spatial size (output = H,W or H/2,W/2 when pooled)
  int relu;
};
// out_lo != null: write fp16 hi / lo planes, C8-planar [image][C/8][H][W][8] (operand split of the tensor-core convs:
// v = hi + lo / 2048); otherwise full fp32 C4-planar
void launch_conv1_direct(LaunchCtx& ctx, const float* img, const float* w9x64, const float* bias,
                         float* out, float* out_lo, int n, int H, int W);
void launch_conv(LaunchCtx& ctx, const ConvParams& p, int ksize, bool pool);
// uint8 grayscale -> fp32 in [0,1]: correctly rounded v / 255 (== float(double(v) / 255), the reference's data loader)
void launch_u8_to_unit_f32(LaunchCtx& ctx, const uint8_t* in, float* out, size_t n);
void launch_c4_to_nchw(LaunchCtx& ctx, const float* in, int c4_total, int c4_off, int C, float* out,
                       int n, int H, int W, bool l2_normalize, const float* in_lo = nullptr);
// full-precision fp32 -> tf32-exact hi / lo planes (element-wise; attention test hook)
void launch_c4_split(LaunchCtx& ctx, const float* in, float* hi, float* lo, size_t n_float4);
// (n,C,H,W) fp32 <-> fp16 hi/lo C8-planar planes (conv layer test hook)
void launch_nchw_to_c8_split(LaunchCtx& ctx, const float* in, int C, void* hi, void* lo, int n, int H, int W);
void launch_c8_to_nchw(LaunchCtx& ctx, const void* hi, const void* lo, int c8_total, int C, float* out, int n, int H, int W);

// ------------------------------------------------------------------ tensor-core 3x3 conv (tc_conv.cu)
struct TcConvParams {
  const float* in_hi; const float* in_lo;   // fp16 C8-planar activation planes (hi, lo*2048), cin/8 units per image
  const float* wpk;                         // tc_conv_pack_weights layout
  const float* bias;                        // [cout_pad]
  float* out_hi; float* out_lo;             // fp16 C8-planar planes; out_lo == null: out_hi = full fp32, C4-planar
  int out_c4_total, out_c4_off;             // channel units per image in the output buffer (8-ch units, or 4-ch if fp32)
  int* overflow;                            // sticky flag: an activation left the fp16 range
  int cin, cout_pad, nb;                    // nb = output channels per CTA tile (64 or 128)
  int n, H, W;
  int relu, pool;
  int ks = 3;                               // 3 (3x3, zero pad 1) or 1 (1x1 heads)
  int in_c8_total = 0, in_c8_off = 0;       // 8-channel units per image in the input buffer (0 = cin/8) / first unit read
  // fused first layer: img != null -> the input activations are computed from the grayscale images (n, H, W) with the
  // stem weights c1 = [9 taps][64] | bias [64] (carried in the kernel parameters) instead of being read from in_hi / in_lo
  const float* img = nullptr;
  const unsigned char* img_u8 = nullptr;    // alternative to img: raw 8-bit pixels, divided by 255 in the patch load
                                            // (datasets/SSHIDataset.py:26-28 + the caller's .float(), correctly rounded)
  float c1[9 * 64 + 64];
  int bias_in_params = 0;                   // cout_pad <= 128: the epilogue reads the bias from bias_c (constant bank)
  float bias_c[128];
  // fused channel-L2 statistics of the descriptor head (superpoint_test.py:125-126): when set (fp32 output only), the
  // epilogue also writes sum_c out[c]^2 over the tile's nb channels per pixel to sumsq[image][cout block][H][W]; the
  // descriptor sampler divides by the norm, so the normalised dense map never exists in HBM
  float* sumsq = nullptr;
  // precision experiment (B200M_SINGLE=desc, profiles/r02_single_product_experiment.md): column blocks >= single_from_cb
  // are computed from the hi planes only (one fp16 product instead of three)
  int single_from_cb = 1 << 30;
};
bool launch_tc_conv(LaunchCtx& ctx, const TcConvParams& p, int num_sms);
size_t tc_conv_weight_floats(int cin, int cout_pad, int nb, int ks);
void tc_conv_pack_weights(const double* w, int cout, int cin, int cout_pad, int nb, int ks, float* dst);
void launch_nchw_to_c4(LaunchCtx& ctx, const float* in, int C, float* out, int c4_total, int n, int H, int W);
// per-pixel sum of squares over C channels of a C4-planar map -> sumsq (n, H, W)  (CUDA-core comparator path only: the
// tensor-core descriptor head produces it in its epilogue)
void launch_c4_sumsq(LaunchCtx& ctx, const float* in, int in_c4_total, int in_c4_off, float* sumsq, int C, int n, int H,
                     int W);

// ------------------------------------------------------------------ detector post (sp_post.cu)
// semi C4-planar (>= 17 groups) -> heat (n, 8h, 8w)
void launch_softmax_heat(LaunchCtx& ctx, const float* semi_c4, int c4_total, float* heat, int n, int hc, int wc);
// heat -> optional dense nms map + candidate list (score, linear index) per image
size_t nms_scratch_bytes(int n, int H8, int W8);
void launch_nms_candidates(LaunchCtx& ctx, const float* heat, float* nms_dense, int n, int H8, int W8,
                           int radius, float thr, int border, unsigned long long* cand_keys,
                           int* cand_counts, int cand_cap, int* overflow_flag, void* scratch);
// candidate list -> keypoints (x,y), scores, counts; descending-score top-k or row-major order
void launch_select_keypoints(LaunchCtx& ctx, unsigned long long* cand_keys, const int* cand_counts,
                             int cand_cap, int n, int W8, int max_kp, float* keypoints, float* scores,
                             int* counts, int cap);
void launch_apply_flags(LaunchCtx& ctx, const int* flags, int* counts, int n);
// C4-planar descriptor map + keypoints -> descriptors (n,D,cap) and/or token-major rows.  sumsq == null: the map is
// already channel-normalised; else the RAW head output with per-pixel partial sums of squares sumsq (n, ncb, hc, wc):
// every tap is divided by its pixel's norm first (same arithmetic as normalising the dense map)
void launch_sample_descriptors(LaunchCtx& ctx, const float* desc_c4, int c4_total, int D, int n, int hc, int wc,
                               const float* keypoints, const int* counts, int cap, int align_corners,
                               float* out_dcn /* (n,D,cap) or null */,
                               float* out_tok /* token-major rows or null */, int tok_ld, size_t tok_img_stride,
                               const float* sumsq = nullptr, int ncb = 1);

// ------------------------------------------------------------------ descriptor 2-NN + ratio test (sp_knn.cu)
// desc (B, D, N) / (B, D, M) channel-major; match (B, N) = nearest train index or -1, dist1/dist2 = Euclidean distances
bool launch_knn_ratio(LaunchCtx& ctx, const float* desc0, const float* desc1, const int* counts0, const int* counts1,
                      int B, int D, int N, int M, float ratio, long long* match, float* dist1, float* dist2);

// ------------------------------------------------------------------ registration after the path (sp_register.cu)
// cv2.estimateAffinePartial2D(RANSAC) over the valid matches of each pair + cv2.warpAffine (superpoint_glue_test.py:83-101)
size_t ransac_smem_bytes(int N);
bool launch_ransac_affine_partial(LaunchCtx& ctx, const float* kpts0, const float* kpts1, const long long* matches0,
                                  const int* counts0, int B, int N, int M, double thr, int max_iters, double confidence,
                                  int refine, double* matrices, unsigned char* inlier0, int* info);
// dtype: 0 = uint8, 1 = float32, 2 = float64
bool launch_warp_affine(LaunchCtx& ctx, const void* src, int dtype, int B, int sH, int sW, const double* matrices,
                        void* dst, int dH, int dW);

// ------------------------------------------------------------------ SuperGlue linear (sg_linear.cu)
struct GemmParams {
  const float* A; int lda; long long strideA;   // [M,K] row-major
  const float* Bw; int ldb; long long strideB;  // [N,K] row-major (weights / second operand)
  float* C; int ldc; long long strideC;         // [M,N]
  const float* bias;                            // [N] or null
  int M, N, K, batch;
  float alpha;       // C = alpha * (A B^T) + bias
  int relu;
  int accumulate;    // C += ...
  // out_f16: C / C_lo / VT / VT_lo are fp16 planes (same element indexing): C = fp16(v), C_lo = fp16(v - C);
  // with VT set, columns >= vt_col0 are ALSO written transposed, VT[row / vt_np][col - vt_col0][row % vt_np]
  int out_f16 = 0;
  float* C_lo = nullptr;
  float* VT = nullptr; float* VT_lo = nullptr; int vt_col0 = 0; int vt_np = 1;
  float lo_scale = 1.f;        // out_f16: C_lo = fp16((v - C) * lo_scale)  (2048 for planes consumed as the weight operand)
  // tensor-core GEMM, batch > 1: rows of A / W of problem b start at b * batch_rows_a / b * batch_rows_b of the same arrays
  int batch_rows_a = 0, batch_rows_b = 0;
  int single = 0;              // precision experiment (B200M_SINGLE=gemm): hi planes only
  // tensor-core GEMM only.  A given as fp16 hi / lo*2048 operand planes [M][lda_p] (halves) instead of fp32: the planes go
  // straight from TMA into the MMA's operand tiles and the in-kernel split (its 64 KB of shared-memory traffic per stage)
  // is skipped.  A stays the fp32 alias for the CUDA-core fallback and may be null when the planes are given.
  const void* A_hi_p = nullptr; const void* A_lo_p = nullptr; int lda_p = 0;
  // fp32 epilogue: ALSO write the result as fp16 hi / lo*2048 planes P_hi / P_lo [M][ldp] (halves) -- the next GEMM's A
  void* P_hi = nullptr; void* P_lo = nullptr; int ldp = 0;
  int stage_planes = 0;        // set by launch_tc_gemm: plane outputs leave through shared-memory staging + TMA tensor stores
};
void launch_gemm(LaunchCtx& ctx, const GemmParams& p);
// tcgen05 fp16x3 version for weight GEMMs (w_hi / w_lo: [N][K] fp16 planes, lo scaled by 2048); false if declined
bool launch_tc_gemm(LaunchCtx& ctx, const GemmParams& p, const float* w_hi, const float* w_lo, int num_sms);
void gemm_pack_fp16_planes(const float* w, size_t n, float* hi_as_float, float* lo_as_float);   // host
// fp32 [rows][ld_in] (first `cols` columns, cols % 8 == 0) -> fp16 operand planes hi / lo*2048 [rows][ld_out] (halves)
void launch_split_planes(LaunchCtx& ctx, const float* in, int ld_in, void* hi, void* lo, int ld_out, size_t rows, int cols);
// (B,C,N) channel-major <-> token-major [B][N][ld] (first C columns)
void launch_bcn_to_tokens(LaunchCtx& ctx, const float* in, int B, int C, int N, float* out, int Np, int ld);
void launch_tokens_to_bcn(LaunchCtx& ctx, const float* in, int Np, int ld, float* out, int B, int C, int N);
// normalised keypoints + score -> [rows][4] (x, y, score, 0)
void launch_kenc_input(LaunchCtx& ctx, const float* kpts, const float* scores, int B, int N, int Np,
                       float cx, float cy, float scale, float* out4);

// ------------------------------------------------------------------ attention (sg_attn.cu)
// qkv: [2*B*Np][3D] head-major columns (q | k | v); cross: source side = other side.
void launch_attention(LaunchCtx& ctx, const float* qkv, float* msg, int B, int Np, int D, int heads,
                      const int* counts0, const int* counts1, int n_full0, int n_full1, bool cross);

// tcgen05 flash attention on tf32 hi/lo planes of the fused q|k|v projection (tc_attn.cu)
// qkv_*: fp16 planes [2*B*Np][3D] (hi = fp16(x), lo = fp16(x - hi)); vt_*: transposed value planes, fp16,
// [2*B blocks][D][Np] (keys contiguous) -- all written by the q|k|v projection's epilogue (GemmParams::out_f16)
// msg_hi / msg_lo set: the message is written as fp16 hi / lo*2048 operand planes [rows][D] instead of fp32 `msg`
bool launch_tc_attention(LaunchCtx& ctx, const void* qkv_hi, const void* qkv_lo, const void* vt_hi,
                         const void* vt_lo, float* msg, int B, int Np, int D,
                         int heads, const int* counts0, const int* counts1, int n_full0, int n_full1, bool cross,
                         void* msg_hi = nullptr, void* msg_lo = nullptr, bool single = false);

// ------------------------------------------------------------------ fused GNN layer tail (tc_gnn.cu), D = 128
// merge -> mlp -> residual -> next layer's q|k|v in one kernel; activations never leave the SM between the GEMMs
struct GnnFusedParams {
  const uint8_t* wts;        // gnn_fused_pack_weights stream of this layer
  const float* bias;         // [256 mlp1 with the merge bias folded in | 128 mlp2 | 384 next q|k|v]
  float* X; int ldx;         // fp32 token state [rows][ldx] (first 128 columns), updated in place
  __half* qkv_hi; __half* qkv_lo;        // next layer's q|k|v planes [rows][384]
  __half* vt_hi; __half* vt_lo; int vt_np;   // transposed V planes [rows / vt_np][128][vt_np]
  int rows;
  int nt4;                   // 3: also produce the next layer's q|k|v; 0: last layer
  int* overflow;             // sticky flag: a state left the fp16 range (or null)
  int single = 0;            // precision experiment (B200M_SINGLE=gnn): hi planes only
};
bool launch_tc_gnn_layer(LaunchCtx& ctx, const GnnFusedParams& p, const void* att_hi, const void* att_lo, int num_sms);
size_t gnn_fused_weight_floats(bool with_qkv);
void gnn_fused_pack_weights(const float* w_merge, const float* w1, const float* w2, const float* w_qkv, float* dst);
// fp32 (rows, 3D) q|k|v buffer -> the fp16 plane set above (test hook; the GEMM epilogue does this in the product path)
void launch_qkv_to_f16_planes(LaunchCtx& ctx, const float* qkv, void* hi, void* lo, void* vt_hi, void* vt_lo,
                              int blocks, int Np, int D);

// ------------------------------------------------------------------ optimal transport (sg_ot.cu)
struct OtParams {
  const float* S; int ldS; long long strideS;   // scores (B, N, M) with leading dim ldS
  float* u; float* v;                           // (B, ld_uv) each
  int ld_uv;
  const int* counts0; const int* counts1;       // device per-pair sizes or null
  int B, N, M;                                  // full sizes (used when counts are null)
  float alpha;
};
void launch_ot_init(LaunchCtx& ctx, const OtParams& p);
void launch_ot_row_update(LaunchCtx& ctx, const OtParams& p);
void launch_ot_col_update(LaunchCtx& ctx, const OtParams& p);
// fused iterations for M <= 1024 (one launch per iteration, S read once per iteration); starts from u = v = 0
bool ot_fused_supported(const OtParams& p);
size_t ot_fused_scratch_floats(int pairs, int N, int M);
void launch_ot_sinkhorn_fused(LaunchCtx& ctx, const OtParams& p, int iters, float* scratch, int num_sms);
void launch_ot_write_Z(LaunchCtx& ctx, const OtParams& p, float* Z);   // dense (B,N+1,M+1), full sizes only
// argmax over rows / columns of the final Z (computed on the fly from S,u,v)
void launch_ot_argmax(LaunchCtx& ctx, const OtParams& p, int* idx0, float* max0, int* idx1);
// same from a dense Z (stage API)
void launch_dense_argmax(LaunchCtx& ctx, const float* Z, int B, int N, int M, int* idx0, float* max0, int* idx1, int ld);
void launch_match_select(LaunchCtx& ctx, const int* idx0, const float* max0, const int* idx1, int ld,
                         const int* counts0, const int* counts1, int B, int N, int M, float thr,
                         long long* matches0, long long* matches1, float* ms0, float* ms1);

// cv2.resize(src, (dw, dh)), uint8 single channel, INTER_LINEAR -- the data loader's resize (datasets/SSHIDataset.py:20-22)
void launch_resize_linear_u8(LaunchCtx& ctx, const unsigned char* src, int B, int sh, int sw, unsigned char* dst, int dh,
                             int dw);
// multi-GPU gather wire format (one int32 buffer per rank: [pair][index row | score-bits row][N])
void launch_pack_match_wire(LaunchCtx& ctx, const long long* matches, const float* scores, int B_valid, int B_wire,
                            int N, int ld, int* wire);
void launch_unpack_match_wire(LaunchCtx& ctx, const int* wire, int world, int bmax, int n_pairs, int N,
                              long long* matches, float* scores);

}  // namespace b200m
