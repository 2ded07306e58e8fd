/*
 * om_b200.h -- C ABI of libom_b200.so: the B200 (sm_100a) implementation of the
 * Shi-Tomasi -> NMS -> top-k -> BAD -> Sinkhorn feature-matching path of
 * fateshelled/onnx_image_processing.
 *
 * The reference has no FFI: its boundary for this path is the Python nn.Module API under
 * pytorch_model/{detector,descriptor,orientation,matching,utils,feature_detection}.  Each entry
 * point below replaces the arithmetic of one of those modules/functions (cited as file:line in
 * the reference checkout); onnx_image_processing_b200/ re-creates the nn.Module surface on top
 * of these calls via ctypes + torch.library custom ops.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name starts with h_; tensors are dense,
 *    row-major, float32 unless stated; images are (B,1,H,W) == (B,H,W).
 *  - keypoints are (B,K,2) float32 in (y,x) order, (-1,-1) for padding entries
 *    (utils/keypoint_utils.py:104-114).
 *  - calls are asynchronous on `stream` (a cudaStream_t passed as void*), never synchronise,
 *    never allocate: scratch memory comes from the caller (`ws`, size from the matching
 *    *_workspace_bytes function).  They are CUDA-graph capturable.
 *  - return value: OM_OK, an OM_ERR_* argument error, or OM_ERR_CUDA_BASE + cudaError_t.
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef OM_B200_H
#define OM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OM_OK 0
#define OM_ERR_NULL 1        /* required pointer is NULL */
#define OM_ERR_SHAPE 2       /* non-positive or inconsistent sizes, K > H*W, ... */
#define OM_ERR_PARAM 3       /* unsupported parameter value (block size, radius, num_pairs, mode) */
#define OM_ERR_WORKSPACE 4   /* ws is NULL or smaller than *_workspace_bytes reports */
#define OM_ERR_LIMIT 5       /* size beyond an implementation limit (documented per call) */
#define OM_ERR_CUDA_BASE 1000

/* descriptor post-processing, descriptor/bad.py:214-218 and :562-567 */
#define OM_DESC_RAW 0        /* centered = diff - threshold */
#define OM_DESC_SOFT 1       /* sigmoid(-centered * temperature) */
#define OM_DESC_HARD 2       /* (centered <= 0) as 0/1 */

/* sampling of the box-average bank, descriptor/bad.py:538-551 */
#define OM_SAMPLE_NEAREST 0
#define OM_SAMPLE_BILINEAR 1

/* where SparseBAD takes the keypoint orientation from */
#define OM_THETA_NONE 0      /* non-oriented, descriptor/bad.py:518-525 */
#define OM_THETA_MAP 1       /* nearest sample of a (B,H,W) orientation map, descriptor/bad.py:487-499 */
#define OM_THETA_MOMENTS 2   /* evaluate orientation/angle_estimation.py:161-170 at the keypoints only */

/* pixel type of the images handed to om_match_pairs / om_detect_u8.  uint8 is an extension for the real callers'
 * preprocessing (sample/visual_odometry.py:65-92, sample/image_matching.py:28-47 produce 8-bit grey images and only then
 * widen them): the kernels read the bytes natively and results are bit-identical to the same pixels passed as float32. */
#define OM_IMAGE_F32 0
#define OM_IMAGE_U8 1

/* matcher flavours of om_match_pairs_f32 */
#define OM_MATCH_SPARSE 0    /* feature_detection/shi_tomasi_sparse_bad_sinkhorn.py:134-182 */
#define OM_MATCH_ANGLE 1     /* feature_detection/shi_tomasi_angle_sparse_bad_sinkhorn.py:132-180 */
#define OM_MATCH_DENSE 2     /* feature_detection/shi_tomasi_bad_sinkhorn.py:162-219 */
#define OM_MATCH_MAPS 3      /* score (and orientation) maps come from the caller's own detector, everything after them is
                              * this library's: feature_detection/akaze_sparse_bad_sinkhorn.py:148-196 (om_match_pairs_from_maps_f32) */

int om_version(void);
/* The library links its own (static) CUDA runtime: select the device the caller's pointers and
 * stream belong to before the first call on a thread (one process per GPU: once, at start-up). */
int om_set_device(int device);
const char* om_error_string(int status);
/* Number of kernels launched by this library since load (all streams); bench.py reads it. */
unsigned long long om_launch_count(void);

/* Host-side copy of the learned BAD tables (descriptor/bad_params.py:4-1568).  h_boxes gets
 * num_pairs*5 signed bytes {x1-16,x2-16,y1-16,y2-16,radius}; h_thresholds num_pairs floats. */
int om_bad_table(int num_pairs, signed char* h_boxes, float* h_thresholds);

/* ---- detector ---------------------------------------------------------------------------- */

/* ShiTomasiScore.forward, detector/shi_tomasi.py:66-112.  block_size odd in [1,9]. */
int om_shi_tomasi_score_f32(const float* image, int B, int H, int W, int block_size,
                            float* score_map, void* stream);

/* apply_nms_maxpool, utils/keypoint_utils.py:12-44.  nms_radius in [0,8]. mask is 0/1 floats. */
int om_nms_mask_f32(const float* scores, int B, int H, int W, int nms_radius,
                    float* mask, void* stream);

size_t om_topk_workspace_bytes(int B, int H, int W, int K);

/* select_topk_keypoints, utils/keypoint_utils.py:47-117, on caller-provided scores and mask.
 * Ties between equal scores go to the lower flat index; output is sorted by descending score.
 * Limit: K <= 16384. */
int om_select_topk_f32(const float* scores, const float* mask, int B, int H, int W, int K,
                       float score_threshold, int border_margin,
                       float* kpts, float* kpt_scores, void* ws, size_t ws_bytes, void* stream);

/* Fused ShiTomasiScore + apply_nms_maxpool + select_topk_keypoints (the three calls at
 * feature_detection/shi_tomasi_sparse_bad_sinkhorn.py:156-173) without materialising the score
 * or mask maps.  score_map may be NULL; if given it receives the full (B,H,W) score map.
 * Workspace: om_topk_workspace_bytes. */
int om_detect_f32(const float* image, int B, int H, int W, int block_size, int nms_radius,
                  int border_margin, float score_threshold, int K,
                  float* score_map, float* kpts, float* kpt_scores,
                  void* ws, size_t ws_bytes, void* stream);

/* Fused apply_nms_maxpool + select_topk_keypoints on a CALLER's score map (utils/keypoint_utils.py:12-117 as called at
 * feature_detection/akaze_sparse_bad_sinkhorn.py:155-170): the NMS mask is never materialised.  scores (B,H,W) >= 0.
 * Workspace: om_topk_workspace_bytes. */
int om_detect_from_scores_f32(const float* scores, int B, int H, int W, int nms_radius, int border_margin,
                              float score_threshold, int K, float* kpts, float* kpt_scores,
                              void* ws, size_t ws_bytes, void* stream);

/* ---- ingest -------------------------------------------------------------------------------- */

/* The real callers' preprocessing in one kernel (sample/visual_odometry.py:65-92 load_image_from_array: cv2.cvtColor(BGR2GRAY),
 * cv2.resize(INTER_LINEAR), astype(float32)): src is (B,Hin,Win,channels) uint8, channels 3 (B,G,R interleaved) or 1; the grey
 * value is OpenCV's 15-bit fixed-point luma, the resize is OpenCV's 11-bit fixed-point bilinear (same coordinates, same
 * coefficient rounding, same border rules, its vector-path final rounding: identical to cv2 4.13 on every size tried).  dst_u8 (B,Hout,Wout) and / or dst_f32 (B,Hout,Wout) receive the
 * result (either may be NULL, not both); the uint8 form feeds om_match_pairs (OM_IMAGE_U8) directly.  Same size in and
 * out copies (cv2.resize does).  No workspace. */
int om_preprocess_u8(const unsigned char* src, int B, int Hin, int Win, int channels, int Hout, int Wout,
                     unsigned char* dst_u8, float* dst_f32, void* stream);

/* ---- orientation ------------------------------------------------------------------------- */

/* AngleEstimator.forward, orientation/angle_estimation.py:123-172.  moment_kernels is the
 * module's (2,1,ps,ps) buffer (x*g, y*g); patch_size odd <= 31.  Zero padding. */
int om_angle_map_f32(const float* image, int B, int H, int W, const float* moment_kernels,
                     int patch_size, float* angle_map, void* stream);

/* ---- descriptors ------------------------------------------------------------------------- */

/* pair_table: (P,6) floats {ox1, ox2, oy1, oy2, radius, threshold}, offsets relative to the
 * keypoint (descriptor/bad.py:405-410), integer-valued, |offset| <= 15, radius in [0,7]. */

size_t om_sparse_bad_workspace_bytes(int B, int H, int W, int theta_mode);

/* SparseBAD.forward, descriptor/bad.py:436-576.  theta_mode selects the oriented branch
 * (:487-517); `orientation` is the (B,H,W) map for OM_THETA_MAP; moment_kernels/patch_size are
 * used by OM_THETA_MOMENTS (same result as running AngleEstimator on the whole image first).
 * The workspace holds the exact integral image of the padded image (built once per call, sampled
 * at the keypoints only) -- instead of the reference's 8-channel box-average bank. */
int om_sparse_bad_f32(const float* image, int B, int H, int W, const float* kpts, int K,
                      const float* pair_table, int P, int desc_mode, float temperature,
                      int normalize, int sampling_mode,
                      int theta_mode, const float* orientation,
                      const float* moment_kernels, int patch_size,
                      float* desc /* B,K,P */, void* ws, size_t ws_bytes, void* stream);

size_t om_dense_bad_workspace_bytes(int B, int H, int W);

/* BADDescriptor.forward (non-oriented), descriptor/bad.py:62-110 and :189-218: the dense
 * (B,P,H,W) map.  Emulates the reference's float32 integral image (double-accumulated cumsum,
 * rounded once per pass) and its 4-tap order. */
int om_dense_bad_f32(const float* image, int B, int H, int W, const float* pair_table, int P,
                     int desc_mode, float temperature, float* desc_map /* B,P,H,W */,
                     void* ws, size_t ws_bytes, void* stream);

/* Dense BAD evaluated only where the dense matcher consumes it: bilinear gather at K keypoints,
 * validity mask and optional L2 normalisation (descriptor/bad.py:277-333,
 * feature_detection/shi_tomasi_bad_sinkhorn.py:120-160 and :212-214). */
int om_dense_bad_at_kpts_f32(const float* image, int B, int H, int W, const float* kpts, int K,
                             const float* pair_table, int P, int desc_mode, float temperature,
                             int normalize, float* desc /* B,K,P */,
                             void* ws, size_t ws_bytes, void* stream);

/* extract_descriptors_at_keypoints (integer gather, descriptor/bad.py:221-274, subpixel=0) and
 * extract_descriptors_at_keypoints_subpixel (bilinear, :277-333, subpixel=1) on a caller map. */
int om_gather_descriptors_f32(const float* desc_map, int B, int D, int H, int W,
                              const float* kpts, int K, int subpixel,
                              float* desc /* B,K,D */, void* stream);

/* ---- matching ---------------------------------------------------------------------------- */

size_t om_sinkhorn_workspace_bytes(int B, int N, int M, int D);

/* SinkhornMatcher.forward, matching/sinkhorn.py:149-208 (cost matrix :79-110, log-domain
 * iterations :112-147).  distance_l1 = 0 -> squared L2, 1 -> L1.  P is (B,N+1,M+1). */
int om_sinkhorn_f32(const float* desc1, const float* desc2, int B, int N, int M, int D,
                    int iterations, float epsilon, float unused_score, int distance_l1,
                    float* P, void* ws, size_t ws_bytes, void* stream);

/* Everything the matching stage can hand back.  When the tcgen05 cluster kernel applies (L2 cost, N, M <= 512, descriptor
 * length a multiple of 32, exp(-unused/eps) not near underflow) the requested pieces are computed in that kernel's epilogue
 * from the probabilities it holds in registers, and `probs` may be NULL: the (N+1)(M+1) matrix then never leaves the chip
 * (what 4 of the reference's 8 exported models need).  Otherwise P is stored (in the workspace if probs is NULL)
 * and the separate kernels below run on it; results are the same.
 *   scores0 / scores1   SinkhornMatcherWithScores, matching/sinkhorn.py:211-259 (NULL: not wanted)
 *   filters             SinkhornMatcherWithFilters, matching/sinkhorn.py:262-465: ratio_threshold <= 0 / dustbin_margin < 0
 *                       disable a test; rejected rows of probs are rewritten (core 0, dustbin 1); filter_valid (B,N) bytes
 *   matches             MutualNearestNeighborMatcher, matching/match_extraction.py:46-184, on the (filtered) matrix:
 *                       matched_kpts1/2 (B,max_matches,2), match_scores (B,max_matches), match_valid (B,max_matches) bytes */
typedef struct om_sinkhorn_outputs {
    float* probs;
    float* scores0;
    float* scores1;
    int filters;
    float ratio_threshold, dustbin_margin;
    unsigned char* filter_valid;
    int matches;
    const float* kpts1;
    const float* kpts2;
    int max_matches;
    float match_threshold;
    float* matched_kpts1;
    float* matched_kpts2;
    float* match_scores;
    unsigned char* match_valid;
} om_sinkhorn_outputs;

size_t om_sinkhorn_ex_workspace_bytes(int B, int N, int M, int D);
int om_sinkhorn_ex_f32(const float* desc1, const float* desc2, int B, int N, int M, int D,
                       int iterations, float epsilon, float unused_score, int distance_l1,
                       const om_sinkhorn_outputs* out, void* ws, size_t ws_bytes, void* stream);

/* SinkhornMatcherWithFilters epilogue, matching/sinkhorn.py:311-465, IN PLACE on probs (B,N+1,M+1): rows that fail
 * the best/second-best ratio test (ratio_threshold <= 0 disables it) or the best-minus-dustbin margin test
 * (dustbin_margin < 0 disables it) get their core zeroed and their dustbin entry set to 1; valid (B,N) bytes 0/1. */
int om_sinkhorn_filter_rows_f32(float* probs, int B, int N, int M, float ratio_threshold, float dustbin_margin,
                                unsigned char* valid, void* stream);

/* SinkhornMatcherWithScores, matching/sinkhorn.py:211-259: scores0 (B,N) = max over the core columns of every row,
 * scores1 (B,M) = max over the core rows of every column of probs (B,N+1,M+1). */
int om_sinkhorn_scores_f32(const float* probs, int B, int N, int M, float* scores0, float* scores1, void* stream);

/* ---- match extraction -------------------------------------------------------------------- */

size_t om_mutual_matches_workspace_bytes(int B, int N, int M);

/* MutualNearestNeighborMatcher.forward, matching/match_extraction.py:46-184 (used behind every matcher by
 * feature_detection/match_extraction_wrapper.py:82-113): mutual row/column argmax of the core block of
 * probs (B,N+1,M+1), probability threshold, the max_matches best in descending order.  Outputs:
 * matched_kpts1/2 (B,max_matches,2), scores (B,max_matches; -1 for rejected rows, 0 for padding beyond N),
 * valid (B,max_matches) bytes 0/1.  Equal scores come out in ascending row order.  Limit: N <= 16384. */
int om_mutual_matches_f32(const float* probs, const float* kpts1, const float* kpts2, int B, int N, int M,
                          int max_matches, float threshold, float* matched_kpts1, float* matched_kpts2,
                          float* scores, unsigned char* valid, void* ws, size_t ws_bytes, void* stream);

/* ---- essential-matrix head --------------------------------------------------------------- */

/* Weighted 8-point essential matrix from the Sinkhorn matrix, one 3x3 per pair (the reference handles one pair per
 * call): EssentialMatrixEstimator.forward, geometry/essential_matrix_estimator.py:292-392 (grid points, valid1 = valid2
 * = NULL, pts_batched = 0: one point set for all pairs) and _estimate_essential_matrix,
 * feature_detection/shi_tomasi_angle_sparse_bad_sinkhorn_essential_matrix.py:184-271 (keypoints, validity masks,
 * pts_batched = 1).  probs (B,N+1,M+1); pts1 (B|1,N,2), pts2 (B|1,M,2): normalised (x, y) = K^-1 [x, y, 1];
 * valid1 (B,N), valid2 (B,M): bytes 0/1, both or neither; E (B,3,3).  Bidirectional top_k mask AND P > 0.01, Hartley
 * normalisation, 9x9 normal equations, n_iter shifted power iterations, denormalisation, projection onto singular
 * values (s, s, 0) with n_iter_manifold power iterations.  One 8-CTA thread-block cluster per pair (row slices, column
 * statistics combined over distributed shared memory).  Limits: top_k <= 8, N, M <= 8192; top_k > N or M is
 * OM_ERR_SHAPE (torch.topk raises). */
int om_essential_matrix_f32(const float* probs, const float* pts1, const float* pts2, const unsigned char* valid1,
                            const unsigned char* valid2, int B, int N, int M, int pts_batched, int top_k, int n_iter,
                            int n_iter_manifold, float* E, void* stream);

/* ---- fused matcher ----------------------------------------------------------------------- */

typedef struct om_match_params {
    int flavour;            /* OM_MATCH_* */
    int B, H, W, K;         /* pairs, image size, max_keypoints */
    int block_size;         /* 3 (sparse/dense default) or 5 (angle default) */
    int nms_radius;
    int border_margin;      /* already resolved (None -> descriptor max radius = 7; dense: 0) */
    float score_threshold;
    int P;                  /* num_pairs: 256 or 512 */
    int desc_mode;          /* OM_DESC_* */
    float temperature;
    int normalize;
    int sampling_mode;      /* OM_SAMPLE_* (sparse/angle only) */
    int patch_size;         /* angle only */
    int iterations;
    float epsilon;
    float unused_score;
    int distance_l1;
    int image_dtype;        /* OM_IMAGE_F32 or OM_IMAGE_U8 */
} om_match_params;

size_t om_match_workspace_bytes(const om_match_params* p);

/* The whole forward of the three unified matcher modules: images in, (kpts1, kpts2, P) out.
 * desc1/desc2 (B,K,P) are optional outputs (NULL: kept in the workspace only). */
int om_match_pairs_f32(const om_match_params* p, const float* image1, const float* image2,
                       const float* pair_table, const float* moment_kernels,
                       float* kpts1, float* kpts2, float* probs /* B,K+1,K+1 */,
                       float* desc1, float* desc2,
                       void* ws, size_t ws_bytes, void* stream);

/* The same with the images in either pixel type (p->image_dtype): image1 / image2 point to float32 or uint8 (B,H,W). */
int om_match_pairs(const om_match_params* p, const void* image1, const void* image2,
                   const float* pair_table, const float* moment_kernels,
                   float* kpts1, float* kpts2, float* probs /* B,K+1,K+1 */,
                   float* desc1, float* desc2,
                   void* ws, size_t ws_bytes, void* stream);

/* The fused matcher with the matching stage's optional outputs (see om_sinkhorn_outputs; out->kpts1/2 are ignored: the
 * matcher's own keypoints are used).  out->probs may be NULL when only matches / scores are wanted.  Workspace:
 * om_match_ex_workspace_bytes. */
size_t om_match_ex_workspace_bytes(const om_match_params* p);
int om_match_pairs_ex(const om_match_params* p, const void* image1, const void* image2,
                      const float* pair_table, const float* moment_kernels,
                      float* kpts1, float* kpts2, float* desc1, float* desc2,
                      const om_sinkhorn_outputs* out, void* ws, size_t ws_bytes, void* stream);

/* OM_MATCH_MAPS: the matcher behind another detector.  scores1/2 (B,H,W) are that detector's score maps, orient1/2 (B,H,W)
 * its orientation maps in radians (both NULL: non-oriented descriptors); NMS + top-k, (oriented) sparse BAD at the keypoints
 * (descriptor/bad.py:487-517 samples the orientation map at the keypoint) and Sinkhorn run here, as in
 * feature_detection/akaze_sparse_bad_sinkhorn.py:148-196.  p->flavour must be OM_MATCH_MAPS, p->image_dtype OM_IMAGE_F32;
 * p->block_size / patch_size are ignored.  Workspace: om_match_workspace_bytes(p). */
int om_match_pairs_from_maps_f32(const om_match_params* p, const float* image1, const float* image2,
                                 const float* scores1, const float* scores2, const float* orient1, const float* orient2,
                                 const float* pair_table, float* kpts1, float* kpts2, float* probs /* B,K+1,K+1 */,
                                 float* desc1, float* desc2, void* ws, size_t ws_bytes, void* stream);

/* om_detect_f32 on uint8 images (block 3 / 5 with NMS radius 3 read the bytes natively, other routings widen first). */
int om_detect_u8(const unsigned char* image, int B, int H, int W, int block_size, int nms_radius,
                 int border_margin, float score_threshold, int K,
                 float* score_map, float* kpts, float* kpt_scores,
                 void* ws, size_t ws_bytes, void* stream);

/* ---- test hooks ---------------------------------------------------------------------------- */

/* Stencil kernel selection: 0 = default routing (block 3 or 5 with radius 3: split register-sweep kernels, i.e. a
 * score kernel and an NMS kernel; block 3 or 5 with radius 5: tiled shared-memory kernel), 1 = generic (runtime block
 * size / radius) kernel, 2 = tiled kernel, 3 = fused sweep kernel, 4 = split sweep kernels; lets the tests check
 * them against each other. */
void om_debug_force_generic_stencil(int on);
/* NMS kernel of the split sweep form at radius 3: 1 = nms3_sweep_kernel (default), 2 = the same at 6 CTAs per SM,
 * 0 = the any-radius nms_sweep_kernel (cross-check). */
void om_debug_nms_variant(int v);
/* Essential-matrix head: 1 = an 8-CTA cluster per pair (default; falls back to one CTA when the per-column scratch does
 * not fit shared memory), 0 = one CTA per pair. */
void om_debug_essential_variant(int clustered);
/* Fused matcher: 4 (default) = image 2's detector / descriptor chain and both integral-image builds run on side streams
 * next to image 1's chain, 2 = only image 2's chain on a side stream, 1 = everything on the caller's stream. */
void om_debug_match_streams(int n);
/* 1 (default): in the fused matcher hard-binarised sparse descriptors (every entry 0 or one value per row) feed the Sinkhorn
 * kernel as ONE 8-bit operand term (popcount similarity on tcgen05 kind::f8f6f4); 0: the two fp16 terms as for any other
 * descriptor. */
void om_debug_match_binary(int on);
/* Keypoint windows of the dense descriptor kernel: bit 0: 1 = one TMA box per keypoint (default), 0 = 16-byte cp.async
 * copies by the keypoint's thread group (cross-check, slower); bit 1 / bit 2 (diagnosis only, wrong results): skip the
 * window fetch / skip the pair arithmetic. */
void om_debug_dense_window(int tma);
/* Banded integral-image build: padded-image rows per band (even, 8..256; default 32).  -16 / -32: the sparse descriptor path
 * stores its integral modulo 2^16 (default) / as uint32. */
void om_debug_band_rows(int rows);
/* Score kernel of the split sweep form: 1 = score3_sweep_kernel / score5_sweep_kernel (default), 2 = the same at
 * another occupancy (6 / 3 CTAs per SM instead of 5 / 4), 0 = stencil_sweep_kernel<.., NMS = false> (cross-check). */
void om_debug_score_variant(int v);
/* Sweep-kernel tuning.  min_blocks 3 or 4: fused sweep kernel, output rows per tile (0 = default 40) and resident
 * CTAs per SM.  min_blocks 99: split kernels, strip_rows for both; 100 + n: strip_rows for the score kernel, n rows
 * for the NMS kernel.  (0, 0) restores every default. */
void om_debug_sweep_tuning(int strip_rows, int min_blocks);

/* Route om_sinkhorn_f32 / the fused matcher through the generic global-memory Sinkhorn kernels
 * instead of the cluster kernel. */
void om_debug_force_generic_sinkhorn(int on);
/* 0: tcgen05/TMEM cluster kernel (default; scaling-form loop when exp(-unused/eps) is safe),
 * 1: FP32-FFMA cluster kernel, 2: generic kernels, 3: tcgen05 kernel with the log-domain loop forced,
 * 4: tcgen05 kernel with the 3xTF32 similarity GEMM forced (default: two-term fp16 split),
 * 5: generic kernels with the log-domain loop forced (2: scaling form when safe),
 * 7: generic kernels with the FP32 FFMA cost GEMM (2 and 5 run the cost GEMM on tcgen05 when D % 32 == 0),
 * 8: the 8-CTA tcgen05 kernel (K <= 512) / generic kernels where 0 takes the hybrid-resident kernel,
 * 9: the streaming kernels of the beyond-1024-keypoint path (sinkhorn_xl.cu) at any size they are eligible for
 *    (query the workspace size with the variant already set). */
void om_debug_sinkhorn_variant(int variant);
/* Streaming path (tuning hook; outputs are bit-identical in every mode): bit 0 set (default 1): odd iterations sweep the
 * matrix backwards, clear: every sweep runs forwards; bit 1 set: plain stream order instead of programmatic dependent
 * launch between its kernels; bits 4..6: L2 eviction policy of the sweeps' bulk copies (0 none, 1 evict-last, 2 / 3 / 4 / 6
 * evict-last on 50 / 75 / 25 / 62.5 % of the lines and evict-first on the rest, 5 evict-first). */
void om_debug_xl_reverse(int on);
/* Device buffer of (B*8 CTAs) x 12 int64: the tcgen05 kernel stores clock64 stamps of its phases there
 * (NULL switches tracing off).  Used by tools/sinkhorn_trace.py only. */
void om_debug_sinkhorn_trace(long long* device_buffer);

/* cudaOccupancyMaxActiveClusters of the hybrid Sinkhorn kernel: big = 0 the 4-CTA form (K <= 512), 1 the 16-CTA form. */
int om_debug_hy_max_clusters(int big);

/* Single-kernel slices of om_detect_f32 / om_dense_bad_at_kpts_f32 so that bench.py can time each
 * kernel with CUDA events.  stage 0 = first kernel(s), stage 1 = the last kernel (needs stage 0's
 * workspace contents). */
int om_debug_detect_stage(const float* image, int B, int H, int W, int block_size, int nms_radius,
                          int border_margin, float score_threshold, int K, float* kpts, float* kpt_scores,
                          void* ws, size_t ws_bytes, void* stream, int stage);
int om_debug_dense_stage(const float* image, int B, int H, int W, const float* kpts, int K,
                         const float* pair_table, int P, int desc_mode, float temperature, int normalize,
                         float* desc, void* ws, size_t ws_bytes, void* stream, int stage);

#ifdef __cplusplus
}
#endif
#endif /* OM_B200_H */
