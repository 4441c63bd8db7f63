/*
 * b200m.h -- C ABI of libb200match.so: the B200-native (sm_100a) SuperPoint + SuperGlue
 * inference hot path of PH8411/image-matching.
 *
 * Boundary being replaced (reference file:line, relative to the reference root):
 *   superglue/models/matching_test.py:54-82      Matching.forward            -> b200m_matching_forward
 *   superpoint/models/superpoint_test.py:103-161 SuperPoint.forward          -> b200m_superpoint_forward
 *   superglue/models/superglue_test.py:230-285   SuperGlue.forward           -> b200m_superglue_forward
 *   superpoint/models/superpoint_test.py:87-100  checkpoint ingestion        -> b200m_set_tensor + b200m_pack
 *   superglue/models/superglue_test.py:221-228   checkpoint ingestion        -> b200m_set_tensor + b200m_pack
 * Stage-level entry points (used by the parity tests, one per reference function):
 *   superpoint_test.py:113-126 (encoder+heads)             -> b200m_superpoint_dense
 *   superpoint_test.py:128-151, 7-37 (softmax/NMS/top-k)   -> b200m_detector_post
 *   superpoint_test.py:40-52   (sample_descriptors)        -> b200m_sample_descriptors
 *   superglue_test.py:63-82, 249-250 (kenc)                -> b200m_keypoint_encode
 *   superglue_test.py:85-138   (AttentionalGNN)            -> b200m_gnn
 *   superglue_test.py:256-260  (final_proj + scores)       -> b200m_score_matrix
 *   superglue_test.py:141-170  (log_optimal_transport)     -> b200m_sinkhorn
 *   superglue_test.py:268-285  (match selection)           -> b200m_match_select
 * Callers' next steps (SURVEY.md 8 f):
 *   superpoint_glue_test.py:83-92 (cv2.estimateAffinePartial2D, RANSAC) -> b200m_estimate_affine_partial
 *   superpoint_glue_test.py:101   (cv2.warpAffine)                      -> b200m_warp_affine
 *   superpoint_flann_test.py:69-78 (FLANN knnMatch + ratio test)        -> b200m_knn_ratio_match
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch types.  All tensor pointers are DEVICE pointers
 *     in the reference's own layouts (NCHW / (B,C,N) fp32, int64 match indices) unless the
 *     name ends in _host.  The library borrows them; it never allocates user-visible memory, and the product entry
 *     points (forward / stage calls listed above) never allocate at all: their scratch is the caller's workspace.  Only
 *     three test hooks -- b200m_debug_conv_layer, b200m_debug_attention and the reference-layout convenience
 *     b200m_sample_descriptors -- take a stream-ordered temporary (cudaMallocAsync / cudaFreeAsync on the call's
 *     stream) for their layout conversion.
 *   - every call takes the CUDA stream to launch on (cudaStream_t passed as void*); no hidden
 *     synchronisation, except where documented (b200m_pack, *_host helpers).
 *   - scratch memory is a caller-owned workspace sized by the matching *_workspace_bytes call.
 *   - return value: 0 = ok, negative = error (B200M_ERR_*); b200m_last_error() returns a
 *     thread-local human-readable message for the last failing call.
 *   - a handle is bound to one device and is not thread-safe (one handle per device/stream).
 */
#ifndef B200M_H
#define B200M_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200M_OK 0
#define B200M_ERR_INVALID (-1)     /* bad argument / unsupported shape            */
#define B200M_ERR_CUDA (-2)        /* a CUDA runtime call or kernel launch failed */
#define B200M_ERR_WORKSPACE (-3)   /* workspace too small                         */
#define B200M_ERR_WEIGHTS (-4)     /* missing / mis-shaped tensor at pack time    */
#define B200M_ERR_RAGGED (-5)      /* per-image keypoint counts differ (reference: torch.stack error,
                                      superglue/models/matching_test.py:75-77)    */

#define B200M_MAX_KENC 8
#define B200M_MAX_GNN 64

typedef struct b200m_handle b200m_handle;

/* Mirrors the reference's config dicts (superpoint_test.py:57-63, superglue_test.py:195-202). */
typedef struct b200m_config {
  int descriptor_dim;            /* D: 64, 128 or 256                                       */
  int nms_radius;                /* reference default 4 (0..4 supported)                     */
  float keypoint_threshold;      /* strict >                                                 */
  int max_keypoints;             /* -1 = keep all (row-major order)                          */
  int remove_borders;            /* reference default 4                                      */
  int align_corners;             /* grid_sample switch the reference derives from the torch
                                    version string (superpoint_test.py:47)                   */
  int n_kenc;                    /* number of hidden keypoint-encoder layers                 */
  int kenc[B200M_MAX_KENC];      /* e.g. {32,64,128}                                         */
  int n_gnn_layers;              /* e.g. 18                                                  */
  int gnn_cross[B200M_MAX_GNN];  /* 0 = 'self', 1 = 'cross'                                  */
  int sinkhorn_iterations;
  float match_threshold;         /* strict >                                                 */
} b200m_config;

const char* b200m_last_error(void);
int b200m_version(void);

int b200m_create(const b200m_config* cfg, int device, b200m_handle** out);
void b200m_destroy(b200m_handle* h);

/* Weight ingestion.  `name` is the reference state_dict key prefixed by "superpoint." or
 * "superglue." (e.g. "superpoint.inc.conv.conv.0.weight", "superglue.gnn.layers.3.attn.proj.1.bias");
 * `data_host` is fp32 host memory, copied.  BatchNorm `num_batches_tracked` is accepted and ignored. */
int b200m_set_tensor(b200m_handle* h, const char* name, const float* data_host,
                     const int64_t* shape, int ndim);
/* Fold BatchNorm, permute attention heads, repack into kernel layouts, upload.  Synchronises `stream`. */
int b200m_pack(b200m_handle* h, void* stream);

/* ---- SuperPoint ---------------------------------------------------------------------------- */
/* Keypoint capacity per image the output buffers must provide: max_keypoints if >= 0, else the
 * candidate capacity for an H x W image. */
int b200m_keypoint_capacity(const b200m_handle* h, int H, int W);
size_t b200m_superpoint_workspace_bytes(const b200m_handle* h, int n_images, int H, int W);
/* images (n,1,H,W) fp32 -> keypoints (n,cap,2) xy fp32, scores (n,cap), descriptors (n,D,cap),
 * counts (n) int32 = valid entries per image (entries beyond count are zero). */
int b200m_superpoint_forward(b200m_handle* h, const float* images, int n_images, int H, int W,
                             float* keypoints, float* scores, float* descriptors, int* counts,
                             int cap, void* ws, size_t ws_bytes, void* stream);
/* uint8 variant, see b200m_matching_forward_u8 */
int b200m_superpoint_forward_u8(b200m_handle* h, const uint8_t* images, int n_images, int H, int W,
                             float* keypoints, float* scores, float* descriptors, int* counts,
                             int cap, void* ws, size_t ws_bytes, void* stream);

/* stage: images -> semi (n,65,h,w), desc (n,D,h,w) channel-L2-normalised (either may be NULL) */
int b200m_superpoint_dense(b200m_handle* h, const float* images, int n_images, int H, int W,
                           float* semi, float* desc, void* ws, size_t ws_bytes, void* stream);
/* stage: semi (n,65,h,w) -> heat (n,8h,8w) [optional], nms (n,8h,8w) [optional],
 *        keypoints (n,cap,2), scores (n,cap), counts (n) */
int b200m_detector_post(b200m_handle* h, const float* semi, int n_images, int hc, int wc,
                        float* heat, float* nms, float* keypoints, float* scores, int* counts,
                        int cap, void* ws, size_t ws_bytes, void* stream);
/* stage: keypoints (n,cap,2), counts (n), desc (n,D,h,w) normalised -> descriptors (n,D,cap) */
int b200m_sample_descriptors(b200m_handle* h, const float* keypoints, const int* counts,
                             const float* desc, int n_images, int hc, int wc, int cap,
                             float* descriptors, void* stream);

/* ---- SuperGlue ----------------------------------------------------------------------------- */
size_t b200m_superglue_workspace_bytes(const b200m_handle* h, int B, int N, int M);
/* keypoints{0,1} (B,N|M,2) xy, scores (B,N|M), descriptors (B,D,N|M); counts{0,1} (B) int32 device
 * arrays of valid entries per pair (NULL = all N / M valid).  Outputs: matches0 (B,N) int64,
 * matches1 (B,M) int64, matching_scores0/1 fp32; entries beyond the counts are -1 / 0. */
int b200m_superglue_forward(b200m_handle* h,
                            const float* kpts0, const float* scores0, const float* desc0, const int* counts0,
                            const float* kpts1, const float* scores1, const float* desc1, const int* counts1,
                            int B, int N, int M, int H0, int W0, int H1, int W1,
                            int64_t* matches0, int64_t* matches1, float* mscores0, float* mscores1,
                            void* ws, size_t ws_bytes, void* stream);

/* stage: -> encoded descriptors (B,D,N) = desc + kenc(normalised kpts, scores) */
int b200m_keypoint_encode(b200m_handle* h, const float* kpts, const float* scores, const float* desc,
                          int B, int N, int H, int W, float* out, void* ws, size_t ws_bytes, void* stream);
/* stage: (B,D,N),(B,D,M) -> (B,D,N),(B,D,M) after layers [layer_begin, layer_end) of the GNN */
int b200m_gnn(b200m_handle* h, const float* desc0, const float* desc1, const int* counts0,
              const int* counts1, int B, int N, int M, int layer_begin, int layer_end,
              float* out0, float* out1, void* ws, size_t ws_bytes, void* stream);
/* stage: final_proj on both sides + scores (B,N,M) */
int b200m_score_matrix(b200m_handle* h, const float* desc0, const float* desc1, int B, int N, int M,
                       float* S, void* ws, size_t ws_bytes, void* stream);
/* stage: S (B,N,M) -> Z (B,N+1,M+1) with the handle's bin_score and `iters` Sinkhorn iterations */
int b200m_sinkhorn(b200m_handle* h, const float* S, int B, int N, int M, int iters, float* Z,
                   void* ws, size_t ws_bytes, void* stream);
/* Descriptor matching after SuperPoint as superpoint_flann_test.py:69-78 does it (cv2 FLANN knnMatch(k=2) + Lowe's
 * ratio test `m.distance < ratio * n.distance`), with an EXACT 2-nearest-neighbour search: desc0 (B,D,N), desc1
 * (B,D,M) in SuperPoint's channel-major layout, optional per-pair counts; matches (B,N) = index of the nearest
 * train descriptor or -1, dist1 / dist2 (B,N) = Euclidean distance to the nearest / second nearest.   */
int b200m_knn_ratio_match(b200m_handle* h, const float* desc0, const float* desc1, const int* counts0,
                          const int* counts1, int B, int N, int M, float ratio, int64_t* matches, float* dist1,
                          float* dist2, void* stream);
/* Registration step of the caller, superpoint_glue_test.py:83-92:
 *   valid = matches > -1; mkpts0 = kpts0[valid]; mkpts1 = kpts1[matches[valid]]
 *   Matrix, mask = cv2.estimateAffinePartial2D(mkpts0, mkpts1, method=cv2.RANSAC, ransacReprojThreshold=thr)
 * for B pairs at once, straight from the device-resident outputs of b200m_matching_forward (no D2H round trip).
 * kpts0 (B,N,2) / kpts1 (B,M,2) xy fp32, matches0 (B,N) int64, counts0 (B) valid keypoints per pair or NULL.
 * Restates OpenCV 4.13's RANSAC (fixed RNG seed, adaptive iteration bound; cv2's defaults are max_iters 2000,
 * confidence 0.99, refine_iters 10): `inlier0` (B,N) uint8 is cv2's mask scattered back to keypoint indices of
 * image0 (bit-identical set), `matrices` (B,2,3) float64 is cv2's matrix to ~1e-12 (the Levenberg-Marquardt
 * refinement is replaced by its closed-form fixed point; refine_iters == 0 skips it like cv2 does).
 * info (B,4) int32: correspondences, inliers, RANSAC iterations run, 1 if a model was found (cv2: Matrix is not None). */
int b200m_estimate_affine_partial(b200m_handle* h, const float* kpts0, const float* kpts1, const int64_t* matches0,
                                  const int* counts0, int B, int N, int M, double ransac_reproj_threshold,
                                  int max_iters, double confidence, int refine_iters, double* matrices,
                                  uint8_t* inlier0, int* info, void* stream);
/* cv2.warpAffine(src, Matrix, (dst_w, dst_h)) with the defaults the caller uses (superpoint_glue_test.py:101:
 * INTER_LINEAR, BORDER_CONSTANT 0, forward matrix inverted internally) for B single-channel images, each with its
 * own matrix (B,2,3) float64 on the device.  dtype: 0 = uint8, 1 = float32, 2 = float64 (the caller's case).
 * Bit-identical to cv2 4.13 (10-bit fixed-point coordinates, 1/32-pixel interpolation grid). */
int b200m_warp_affine(b200m_handle* h, const void* src, int dtype, int B, int src_h, int src_w,
                      const double* matrices, void* dst, int dst_h, int dst_w, void* stream);
/* stage: Z (B,N+1,M+1) -> matches / scores */
int b200m_match_select(b200m_handle* h, const float* Z, int B, int N, int M,
                       int64_t* matches0, int64_t* matches1, float* mscores0, float* mscores1,
                       void* ws, size_t ws_bytes, void* stream);

/* ---- Matching (the drop-in entry point) ---------------------------------------------------- */
/* Workspace for B pairs.  For B <= 16 it holds one SuperPoint workspace PER IMAGE SIDE: the two sides' SuperPoint
 * passes of a small batch then run concurrently (image1 on a handle-owned stream forked from / joined to `stream`,
 * a parallel branch of the captured graph) -- the reference's caller runs one pair per call
 * (superpoint_glue_test.py:65-78), where a single image leaves most of the chip idle.  A smaller workspace (the
 * single-side size) is accepted and runs the sides one after the other.  B200M_SP_DUAL_MAX overrides the bound. */
size_t b200m_matching_workspace_bytes(const b200m_handle* h, int B, int H, int W);
/* image0/image1 (B,1,H,W) fp32 device.  SuperPoint outputs as in b200m_superpoint_forward for each
 * side (capacity `cap`); matches0/1 (B,cap) int64, matching_scores0/1 (B,cap).  No host sync:
 * ragged per-pair keypoint counts are handled on the device (counts0/1 tell the caller how many
 * leading entries are valid); the Python shim reproduces the reference's torch.stack error. */
int b200m_matching_forward(b200m_handle* h, const float* image0, const float* image1, int B, int H, int W,
                           float* keypoints0, float* scores0, float* descriptors0, int* counts0,
                           float* keypoints1, float* scores1, float* descriptors1, int* counts1,
                           int cap, int64_t* matches0, int64_t* matches1, float* mscores0, float* mscores1,
                           void* ws, size_t ws_bytes, void* stream);

/* Same with 8-bit grayscale images (B,1,H,W) uint8 on the device: the data loader's `img / 255.` normalisation
 * (datasets/SSHIDataset.py:26-29 followed by the caller's `.float()`, superpoint_glue_test.py:74-75) is done on the
 * device, correctly rounded, so the results are bit-identical to uploading the fp32 tensor -- with 4x fewer bytes
 * over PCIe.   */
int b200m_matching_forward_u8(b200m_handle* h, const uint8_t* image0, const uint8_t* image1, int B, int H, int W,
                           float* keypoints0, float* scores0, float* descriptors0, int* counts0,
                           float* keypoints1, float* scores1, float* descriptors1, int* counts1,
                           int cap, int64_t* matches0, int64_t* matches1, float* mscores0, float* mscores1,
                           void* ws, size_t ws_bytes, void* stream);

/* The data loader's resize: cv2.resize(img, (dst_w, dst_h)) of datasets/SSHIDataset.py:20-22 -- 8-bit
 * grayscale, default INTER_LINEAR -- for B images (B,src_h,src_w) uint8 on the device, bit-identical to OpenCV 4.13
 * (11-bit fixed-point bilinear; its 2x2 box filter for an exact 2x decimation).  The caller passes
 * dst_w = int(resize_scale * src_w), dst_h = int(resize_scale * src_h) like the reference does.  The result feeds
 * b200m_matching_forward_u8 / b200m_superpoint_forward_u8, whose first kernel divides by 255 while it loads the pixels. */
int b200m_resize_linear_u8(b200m_handle* h, const uint8_t* src, int B, int src_h, int src_w, uint8_t* dst, int dst_h,
                           int dst_w, void* stream);

/* Multi-GPU gather of the results (pairs are sharded over ranks, the only collective is the final gather
 * of match indices): ONE int32 wire buffer per rank, (B_wire, 2, N) = per pair a row of match indices (int64 -> int32)
 * and a row of matching-score bits, so a single all-gather moves both.  `pack` reads this rank's (B_valid, N) results
 * with row pitch `ld` and pads pairs [B_valid, B_wire) with -1 / 0; `unpack` turns the gathered (world, B_wire, 2, N)
 * buffer into (n_pairs, N) int64 matches + fp32 scores (contiguous shards, remainder pairs on the first ranks). */
int b200m_pack_match_wire(b200m_handle* h, const int64_t* matches0, const float* mscores0, int B_valid, int B_wire,
                          int N, int ld, int32_t* wire, void* stream);
int b200m_unpack_match_wire(b200m_handle* h, const int32_t* wire, int world, int B_wire, int n_pairs, int N,
                            int64_t* matches0, float* mscores0, void* stream);

/* Test hook: run ONE packed 3x3 encoder/head layer (0=inc.conv[3], 1..2=down1, 3..4=down2, 5..6=down3, 7=convPa|convDa)
 * on `in` (n,cin,H,W) -> `out` (n,cout,H',W') with the tcgen05 fp16 hi/lo-split kernel (use_tc=1) or the fp32 CUDA-core
 * kernel (use_tc=0), so the two implementations can be compared layer by layer. */
int b200m_debug_conv_layer(b200m_handle* h, int layer, int use_tc, const float* in, float* out, int n, int H,
                           int W, void* stream);

/* Test hook: multi-head attention on a fused projection buffer qkv (2*B*Np rows x 3D, head-major q|k|v columns)
 * -> msg (2*B*Np x D); n0/n1 valid tokens per side; tcgen05 kernel (use_tc=1) or fp32 CUDA-core kernel (0). */
int b200m_debug_attention(b200m_handle* h, const float* qkv, float* msg, int B, int Np, int n0, int n1, int cross,
                          int use_tc, void* stream);

/* Number of kernels launched by this handle since creation (bench.py's gpu_launches); kernels executed through a
 * replayed CUDA graph are counted like directly launched ones. */
long long b200m_launch_count(const b200m_handle* h);
/* The forward entry points (b200m_matching_forward[_u8], b200m_superpoint_forward[_u8], b200m_superglue_forward) capture
 * their launch sequence into a CUDA graph the second time they are called with identical arguments (same shapes,
 * pointers and workspace) and replay it afterwards; semantics are unchanged (stream-ordered, caller-owned buffers).
 * B200M_GRAPHS=0 in the environment disables it.  This returns how many calls were served by a replay. */
long long b200m_graph_replay_count(const b200m_handle* h);

/* Per-launch CUDA-event timing on the launching stream (bench.py's roofline leg; off by default).
 * b200m_profile_end synchronises the device and writes {"kernel": {"ms": total, "launches": n}, ...}. */
int b200m_profile_begin(b200m_handle* h, int max_records);
int b200m_profile_end(b200m_handle* h, char* json, size_t json_cap);

#ifdef __cplusplus
}
#endif
#endif /* B200M_H */
