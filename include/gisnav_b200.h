/*
 * gisnav_b200.h — C ABI of libgisnav_b200.so, the B200 (sm_100a) implementation of GISNav's
 * pose-estimation hot path.  Plain pointers and sizes only; no torch / CUDA types in signatures
 * (a cudaStream_t travels as void*).  Citations are relative to the reference tree.
 *
 * The reference has no plugin ABC for this path; the boundary is the three
 * call sites PoseNode uses, and each entry point below names the one it replaces:
 *
 *   gnb_extract          <- self._extractor.detectAndCompute(ref, None)
 *                            ros/gisnav/gisnav/core/pose_node.py:230 (also twist_node.py:227,230)
 *   gnb_match            <- self._matcher(descs_qry, descs_ref, lafs_qry, lafs_ref)
 *                            pose_node.py:285-287 (+ RootSIFT prep :279-284, gather :296-297)
 *   gnb_solve_pnp        <- compute_pose(camera_info, mkp_qry, mkp_ref, elevation)
 *                            ros/gisnav/gisnav/core/_shared.py:89-125
 *   gnb_geodetic_tail    <- pose_node.py:333-381 with _transformations.py:301-327,330-346,369-393
 *   gnb_pose_batch       <- the whole of PoseNode._pose lines 226-381 for B independent
 *                            (query frame, map tile) pairs, everything resident on the device
 *
 * Status convention (reference: "cannot compute" => return None, pose_node.py:299-307;
 * exceptions are never caught inside _pose): 0 = ok; >0 = soft failure the Python wrapper maps to
 * None; <0 = usage or CUDA error (wrapper raises).
 *
 * Threading (reference: MultiThreadedExecutor, callbacks mutually exclusive,
 * ros/gisnav/gisnav/__init__.py:140-154): one call in flight per gnb_ctx; a ctx may be used from
 * any host thread (each entry point sets the device).  Input buffers are never retained.
 */
#ifndef GISNAV_B200_H
#define GISNAV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNB_DESC_DIM 256

/* status codes */
#define GNB_OK 0
#define GNB_SOFT_TOO_FEW_MATCHES 1 /* < min_matches (pose_node.py:63,299-303) */
#define GNB_SOFT_PNP_FAILED 2      /* no hypothesis with >= 4 inliers */
#define GNB_SOFT_OUT_OF_BOUNDS 3   /* camera centre outside the raster (pose_node.py:340-342) */
#define GNB_E_INVALID (-1)         /* bad argument */
#define GNB_E_CUDA (-2)            /* CUDA runtime error; see gnb_last_error */
#define GNB_E_CAPACITY (-3)        /* exceeds the capacity the ctx was created with */
#define GNB_E_RANGE (-4)           /* reference keypoint outside the DEM (numpy IndexError, _shared.py:100-101) */
#define GNB_E_NO_DEVICE (-5)       /* no sm_100 device / driver */

typedef struct gnb_ctx gnb_ctx;

/* Plain config struct (the reference has class constants instead of parameters:
 * pose_node.py:60-72, launch/params/pose_node.yaml:1-3). */
typedef struct gnb_config {
    int32_t max_keypoints;      /* K cap per image (reference constant 1024, pose_node.py:66) */
    int32_t nms_radius;         /* 4 */
    float keypoint_threshold;   /* 0.005 */
    int32_t border;             /* 4 */
    float match_threshold;      /* CONFIDENCE_THRESHOLD = 0.5, pose_node.py:60 */
    int32_t min_matches;        /* MIN_MATCHES = 15, pose_node.py:63 */
    int32_t ransac_iters;       /* fixed hypothesis count (reference: iterationsCount=10 with
                                   confidence early-exit, _shared.py:115) */
    float reproj_px;            /* 8.0, OpenCV default used by the reference */
    uint32_t ransac_seed;
    int32_t refine;             /* 1 = LM refit on the winner's inliers */
    int32_t max_batch;          /* pairs per gnb_pose_batch call */
    int32_t max_image_h;        /* largest image side the workspace is sized for */
    int32_t max_image_w;
    int32_t conv_impl;          /* 0 = tcgen05 implicit GEMM (product path); 1 = SIMT validation kernel */
    int32_t match_impl;         /* 0 = tcgen05 descriptor GEMM (product path); 1 = SIMT validation kernel */
    int32_t tile_cache;         /* reference-raster feature cache entries (>= max_batch; pose_node.py:226-241) */
    int32_t precision;          /* arithmetic of the dense stack and the matcher head.
                                   0 = fast: bf16 operands, fp32 accumulation (activations stored as bf16);
                                   1 = fp32-faithful: every conv operand is split into two bf16 terms (v = hi + lo,
                                       16 significant bits) and three tcgen05 MMAs (hi*hi + hi*lo + lo*hi) accumulate
                                       into one fp32 TMEM accumulator; conv1a, the two 1x1 heads and the matcher head
                                       run in plain fp32.  Reproduces the reference's fp32 tensors
                                       (pose_node.py:254-287) to ~1e-5 relative; about 3x the tensor work. */
} gnb_config;

/* Fill cfg with the defaults listed above (K=1024, iters=2048, batch 8, 1088x1280 workspace). */
int gnb_default_config(gnb_config* cfg);

/* Build a context on CUDA device `device`: allocates the workspace, repacks the weight blob
 * (layout: gisnav_b200/weights.py) into per-tap bf16 tiles ON THE DEVICE.  weights may be a host pointer
 * or, after an NCCL broadcast, a device pointer (weights_on_device != 0: consumed in place, no host copy). */
int gnb_create(const gnb_config* cfg, const void* weights, size_t nbytes, int weights_on_device,
               int device, gnb_ctx** out);
void gnb_destroy(gnb_ctx* ctx);
const char* gnb_last_error(const gnb_ctx* ctx); /* never NULL; ctx may be NULL for create errors */
int gnb_get_config(const gnb_ctx* ctx, gnb_config* out);
/* Number of kernels this library has launched on ctx since creation (bench "gpu_launches"). */
int64_t gnb_launch_count(const gnb_ctx* ctx);
/* Stream all work of this ctx is issued on (a cudaStream_t), for event timing by the caller. */
void* gnb_stream(const gnb_ctx* ctx);
/* Per-kernel CUDA-event timing on that stream: while enabled, every kernel launch is bracketed by an
 * event pair; gnb_profile_read synchronises, aggregates by kernel name (names: cap x 64 chars,
 * total_ms and launches: cap entries) and clears the log. */
int gnb_profile_enable(gnb_ctx* ctx, int on);
int gnb_profile_read(gnb_ctx* ctx, char* names, float* total_ms, int64_t* launches, int cap, int* n_out);

/* ---- the three reference call sites ------------------------------------------------------- */

/* detectAndCompute: image u8 [h, w] with row stride `stride` bytes (h, w multiples of 8) ->
 * up to `cap` keypoints sorted by descending score (ties: ascending y*w+x):
 * out_xy f32 [cap,2] (x=col, y=row), out_score f32 [cap], out_desc f32 [cap,256]. */
int gnb_extract(gnb_ctx* ctx, const uint8_t* image, int h, int w, int stride, int on_device,
                float* out_xy, float* out_score, float* out_desc, int cap, int* n_out);

/* matcher(desc1, desc2, ...): desc_a f32 [n_a,256], desc_b f32 [n_b,256] ->
 * out_idx int64 [cap,2] (col 0 indexes desc_a, col 1 desc_b; rows sorted by col 0),
 * out_score f32 [cap] = exp(assignment log-score) > match_threshold. */
int gnb_match(gnb_ctx* ctx, const float* desc_a, int n_a, const float* desc_b, int n_b, int on_device,
              int64_t* out_idx, float* out_score, int cap, int* n_out);

/* The reference matcher's transformer layers — LightGlueMatcher("sift", {"n_layers": 9, "depth_confidence": -1,
 * "width_confidence": -1, ...}), pose_node.py:109-121 — are optional here: load a GNBL layer blob
 * (gisnav_b200/weights.py: pack_layers) and every matcher call runs n_layers of self + cross attention in front
 * of the assignment head (no early exit, no pruning, like the reference's active branch).  blob NULL / nbytes 0
 * unloads them.  Needs max_keypoints % 8 == 0 and match_impl = 0. */
int gnb_set_matcher_layers(gnb_ctx* ctx, const void* blob, size_t nbytes);
int gnb_matcher_layers(const gnb_ctx* ctx); /* number of loaded layers, 0 = head only */

/* matcher(desc1, desc2, lafs1, lafs2) with the keypoint centres of the LAFs (pose_node.py:267-276,285-287):
 * kp f32 [n,2] pixel (x, y); (h, w) = image size used to normalise them (the reference passes none and kornia
 * then uses the largest keypoint coordinate per axis — the Python wrapper reproduces that).  Same outputs as
 * gnb_match.  Works without loaded layers too (keypoints are then unused). */
int gnb_match_lightglue(gnb_ctx* ctx, const float* desc_a, const float* kp_a, int n_a, float h_a, float w_a,
                        const float* desc_b, const float* kp_b, int n_b, float h_b, float w_b, int on_device,
                        int64_t* out_idx, float* out_score, int cap, int* n_out);

/* TwistNode's visual-odometry matcher — self._bf.knnMatch(desc_qry, desc_ref, k=2) + ratio test
 * `m.distance < 0.7 * n.distance` (ros/gisnav/gisnav/core/twist_node.py:95,248,263-267).  desc f32
 * [n,dim], dim <= 256 (SIFT: 128); buffers all host or all device (on_device).  out_idx int64 [cap,2]
 * (queryIdx, trainIdx) in query order, out_dist f32 [cap] = m.distance.  `ratio` is a double and the test
 * is evaluated in float64 like the Python expression.  Exact (bit-identical to OpenCV) for integer-valued
 * descriptors 0..255 such as SIFT's; other values are rounded to bf16 (the tensor-core operand type). */
int gnb_knn_ratio_match(gnb_ctx* ctx, const float* desc_q, int n_q, const float* desc_r, int n_r, int dim, double ratio,
                        int on_device, int64_t* out_idx, float* out_dist, int cap, int* n_out);

/* compute_pose: mkp_qry f32 [n,2], mkp_ref f32 [n,2], elevation u8 [dem_h,dem_w] (NULL => z=0),
 * k f64 [9] row-major -> r f64 [9] row-major, t f64 [3]; optional inlier mask u8 [n]. */
int gnb_solve_pnp(gnb_ctx* ctx, const float* mkp_qry, const float* mkp_ref, int n, const uint8_t* dem,
                  int dem_h, int dem_w, const double* k9, int on_device, double* out_r9, double* out_t3,
                  uint8_t* out_inlier_mask, int* n_inliers);

/* pose tail on the host-visible result: r9,t3 from gnb_solve_pnp, affine f64 [12] = 3x4 row-major
 * pixel->(lon,lat,alt) (+proj=affine).  -> ecef f64 [3], quat f64 [4] (x,y,z,w), lla f64 [3]. */
int gnb_geodetic_tail(gnb_ctx* ctx, const double* r9, const double* t3, const double* affine12, int ref_h,
                      int ref_w, double* out_ecef3, double* out_quat4, double* out_lla3);

/* ---- fused, batched path ------------------------------------------------------------------ */

typedef struct gnb_pose_result {
    int32_t status;     /* GNB_OK or GNB_SOFT_* per pair */
    int32_t n_kp_qry, n_kp_ref, n_matches, n_inliers;
    int32_t best_hypothesis;
    double r[9];        /* raster frame -> camera */
    double t[3];
    double ecef[3];     /* camera centre, metres */
    double quat[4];     /* camera orientation in ECEF, x,y,z,w */
    double lla[3];      /* lon, lat, alt */
} gnb_pose_result;

/* B pairs: frames u8 [B,hq,wq], tiles u8 [B,ht,wt], dems u8 [B,ht,wt] (NULL => zeros),
 * k9 f64 [B,9], affine12 f64 [B,12].  All inputs host or all device (on_device).
 * results: HOST array of B gnb_pose_result. */
int gnb_pose_batch(gnb_ctx* ctx, int batch, const uint8_t* frames, int hq, int wq, const uint8_t* tiles,
                   int ht, int wt, const uint8_t* dems, const double* k9, const double* affine12,
                   int on_device, gnb_pose_result* results);

/* PoseNode._pose as the reference receives its inputs (OrthoStereoImage, ros/gisnav_msgs/msg/OrthoStereoImage.msg:14-18):
 * the query side is the PointCloud2 `query_sift` payload — n_records packed keypoint records of point_step bytes
 * (x, y, z, size, angle f32 + descriptor f32[desc_dim]; KEYPOINT_DTYPE, _shared.py:26-35, pose_node.py:207-213) — and
 * the reference side the mono8 raster + DEM.  records: HOST bytes, copied as they are and unpacked on the device.
 * desc_dim must be 256 (point_step 1044): the 128-d SIFT layout cannot feed this matcher head.  (hq, wq) = size of the
 * image the query keypoints came from (used by the transformer layers' position encoding; 0 = raster size).
 * result: one HOST gnb_pose_result. */
int gnb_pose_from_records(gnb_ctx* ctx, const void* records, int n_records, int point_step, int desc_dim, int hq, int wq,
                          const uint8_t* reference, int ht, int wt, const uint8_t* dem, const double* k9,
                          const double* affine12, gnb_pose_result* result);

/* Candidate search for ONE query frame (BASELINE.json config 4): the frame is extracted once and
 * matched against n_tiles reference rasters (n_tiles <= max_batch).  Raster features are cached on
 * the device keyed by tile_ids[i] (the reference re-extracts a raster only when its stamp changes,
 * pose_node.py:226-241); id < 0 = never cached; tile_ids NULL = no caching.  tiles u8 [n,ht,wt],
 * dems u8 [n,ht,wt] or NULL, k9 f64 [9], affine12 f64 [n,12].  results: HOST array of n_tiles;
 * n_cache_hits (optional) = rasters whose features came from the cache.  With transformer layers loaded the cache
 * keeps the rasters' raw keypoints + descriptors and the frame is refined against every candidate separately. */
int gnb_pose_candidates(gnb_ctx* ctx, const uint8_t* frame, int hq, int wq, int n_tiles, const uint8_t* tiles, int ht,
                        int wt, const int64_t* tile_ids, const uint8_t* dems, const double* k9, const double* affine12,
                        gnb_pose_result* results, int* n_cache_hits);
/* The same with one HOST pointer per raster (u8 [ht,wt] each, contiguous).  The pointer of a raster that the cache will
 * serve may be NULL: gnb_cache_lookup says which (hit_out[i] = 1), so the caller of a flyover stream gathers pixels only
 * for the rasters that are new (pose_node.py:226-241 re-extracts a raster only when its stamp changes).  A NULL pointer for
 * a raster that is NOT cached -> GNB_E_INVALID, nothing modified. */
int gnb_pose_candidates_ptrs(gnb_ctx* ctx, const uint8_t* frame, int hq, int wq, int n_tiles, const uint8_t* const* tile_ptrs,
                             int ht, int wt, const int64_t* tile_ids, const uint8_t* dems, const double* k9,
                             const double* affine12, gnb_pose_result* results, int* n_cache_hits);
int gnb_cache_lookup(gnb_ctx* ctx, const int64_t* tile_ids, int n_tiles, int ht, int wt, int* hit_out);
int gnb_cache_clear(gnb_ctx* ctx);

/* ---- the step in front of the path: StereoNode's rotate + centre-crop ------ */

/* StereoNode._rotate_and_crop_center (ros/gisnav/gisnav/core/stereo_node.py:292-335) with the
 * grayscale conversion in front of it (cv2.cvtColor BGR2GRAY, stereo_node.py:239) fused in:
 * ortho u8 [h,w,channels] (channels 1 = gray, 3 = BGR interleaved), dem u8 [h,w] or NULL ->
 * out_ref u8 [crop_h,crop_w], out_dem u8 [crop_h,crop_w] = cv2.warpAffine(stack,
 * getRotationMatrix2D((w//2,h//2), angle, 1), (w,h))[dy:dy+crop_h, dx:dx+crop_w], bit-identical to
 * OpenCV's fixed-point INTER_LINEAR / BORDER_CONSTANT path.  out_rotation6 (optional) = the 2x3
 * rotation matrix; out_inverse9 (optional) = the 3x3 matrix mapping cropped-frame pixels back to
 * the original raster (second return value of the reference function).  Buffers all host or all
 * device (on_device). */
int gnb_rotate_crop(gnb_ctx* ctx, const uint8_t* ortho, int channels, const uint8_t* dem, int h, int w,
                    double angle_degrees, int crop_h, int crop_w, int on_device, uint8_t* out_ref, uint8_t* out_dem,
                    double* out_rotation6, double* out_inverse9);

/* ---- stage-isolated hooks (parity tests feed the oracle's intermediate into one stage) ------ */

/* K1: image -> score map f32 [h,w] and L2-normalised dense descriptors f32 [h/8,w/8,256]. */
int gnb_dense(gnb_ctx* ctx, const uint8_t* image, int h, int w, int stride, float* out_score,
              float* out_dense);
/* K1 per-layer taps: layer in {"conv1a","pool1","conv2a","pool2","conv3a","pool3","conv4a","conv4b",
 * "convPa","convDa"} after running gnb_dense; out f32 [hl,wl,c] (converted from bf16). */
int gnb_layer_activation(gnb_ctx* ctx, const char* layer, float* out, size_t out_floats);
/* K1 on a batch of n images u8 [n,h,w] (host), n <= max_batch: the dense stack, optionally the dense
 * descriptor map (dense_desc) and K2 + K3 into keypoint slots [0, n) (with_keypoints).  Parity hook for
 * the persistent multi-tile, multi-image loops of the conv kernels at BASELINE.json's full sizes. */
int gnb_dense_batch(gnb_ctx* ctx, const uint8_t* images, int n, int h, int w, int dense_desc, int with_keypoints);
/* gnb_layer_activation for image `image_index` of the last pass; also accepts "score" (f32 [h,w]) and
 * "dense" (f32 [h/8,w/8,256], only after a pass with dense_desc). */
int gnb_layer_activation_at(gnb_ctx* ctx, const char* layer, int image_index, float* out, size_t out_floats);
/* keypoints (x, y) f32 [n,2], scores f32 [n] and descriptors f32 [n,256] held in keypoint slot `slot`
 * (slots [0, max_batch) = query frames / gnb_dense_batch images, [max_batch, 2 max_batch) = rasters). */
int gnb_slot_keypoints(gnb_ctx* ctx, int slot, float* out_xy, float* out_score, float* out_desc, int cap, int* n_out);
/* match index pairs int32 [n,2] (query keypoint, reference keypoint; rows sorted by column 0) of pair `pair` of
 * the last gnb_pose_batch / gnb_pose_candidates / matcher call. */
int gnb_pair_matches(gnb_ctx* ctx, int pair, int32_t* out_idx, int cap, int* n_out);
/* K2: NMS + threshold + border + top-K on a caller-supplied score map. */
int gnb_select_keypoints(gnb_ctx* ctx, const float* score, int h, int w, float* out_xy, float* out_score,
                         int cap, int* n_out);
/* K3: bilinear descriptor sampling + L2 norm on a caller-supplied dense map f32 [hc,wc,256]. */
int gnb_sample_descriptors(gnb_ctx* ctx, const float* dense, int hc, int wc, const float* xy, int n,
                           int img_h, int img_w, float* out_desc);
/* K4 internals: full assignment log-score matrix f32 [n_a,n_b] (small sizes only). */
int gnb_match_scores(gnb_ctx* ctx, const float* desc_a, int n_a, const float* desc_b, int n_b,
                     float* out_scores);
/* transformer internals: descriptors f32 [n,256] of side 0 / 1 after the layers of the last matcher call. */
int gnb_refined_descriptors(gnb_ctx* ctx, int side, float* out, int n);
/* K5 internals: per-hypothesis inlier counts int32 [ransac_iters] (-1 = invalid hypothesis) and
 * hypotheses f32 [ransac_iters,12] (R row-major, t) for the last gnb_solve_pnp call. */
int gnb_ransac_debug(gnb_ctx* ctx, int32_t* out_counts, float* out_hyp, int* out_best);

#ifdef __cplusplus
}
#endif
#endif /* GISNAV_B200_H */
