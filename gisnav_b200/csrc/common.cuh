// common.cuh — context, workspace and launch helpers shared by the sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gisnav_b200.h"

typedef __nv_bfloat16 bf16;

#define GNB_NUM_LAYERS 12
#define GNB_CAND_CAP 65536  // NMS survivors kept per image before top-K
#define GNB_NMS_LIST_CAP 131072   // listed pixels per image of the list-based NMS (more: that image falls back to the tile kernel)
#define GNB_MAX_KP 4096     // hard ceiling for cfg.max_keypoints

enum LayerId { L1A = 0, L1B, L2A, L2B, L3A, L3B, L4A, L4B, LPA, LPB, LDA, LDB };

struct ConvLayer {
    int cin, cout, ks, cout_pad;
    bf16* w;      // device, [ks*ks][cout_pad][cin]  (K-major per tap: tcgen05 B operand / SIMT)
    float* bias;  // device, [cout_pad]
    // fp32-faithful mode (cfg.precision = 1) only:
    bf16* w_x3;   // device, [ks*ks][cout_pad][2 cin]: per output row the hi terms of its cin weights, then the lo terms
    float* w_f32; // device, [ks*ks][cout_pad][cin] fp32 (conv1a and the two 1x1 heads run on the CUDA cores in fp32)
};

// Per-image-slot activation pointers for one batched pass of the conv stack (all NHWC).
struct ConvWorkspace {
    int cap_images;       // images per pass
    size_t cap_pixels;    // max h*w per image
    uint8_t* img;         // [n][h][w]  input of the current pass (points at img_a or img_b)
    uint8_t *img_a, *img_b; // two staging buffers so the second H2D copy overlaps the first pass
    bf16 *a1a, *p1, *a2a, *p2, *a3a, *p3, *a4a, *a4b, *apa, *ada;
    float* semi;          // [n][hc][wc][65]
    float* score;         // [n][h][w]
    float* dense;         // [n][hc][wc][256]  L2-normalised
    int n, h, w;          // shape of the last pass (for gnb_layer_activation)
};

// Device-side mirror of gnb_pose_result with the fields the kernels fill.
struct PairOut {
    int32_t status, n_kp_qry, n_kp_ref, n_matches, n_inliers, best_hypothesis, pad0, pad1;
    double r[9], t[3], ecef[3], quat[4], lla[3];
};

struct gnb_ctx;
int gnb_tc_err_check(gnb_ctx* ctx);  // after a stream sync: did a bounded tcgen05 pipeline wait time out?

struct gnb_ctx {
    gnb_config cfg;
    int device;
    int sm_count;
    cudaStream_t stream;
    cudaStream_t copy_stream;           // H2D staging of the batch path, overlapped with compute
    cudaStream_t stream2;               // small batches: the raster chain (K1-K3) runs here, concurrently with the frame chain
    cudaEvent_t ev_join;
    cudaEvent_t ev_frames, ev_tiles, ev_params;
    char err[512];
    int64_t launches;
    ConvLayer layers[GNB_NUM_LAYERS];
    // matcher head
    bf16* match_w;    // [256 out][256 in]
    bf16* match_wt;   // [256 in][256 out] transposed copy for the SIMT projection
    float* match_b;   // [256]
    bf16* match_mw;   // [256]
    float match_mb;
    ConvWorkspace cw;
    ConvWorkspace cw2;                  // second set of activation buffers for `overlap_images` images (0: none), see stream2
    int overlap_images;
    // keypoints: slots 0..2*max_batch-1 (frames then tiles)
    int kp_slots;
    unsigned long long* cand_keys;  // [slots][GNB_CAND_CAP]
    int* cand_count;                // [slots]
    unsigned* nms_hist;             // [slots][2048] score-bit histogram of the sparse NMS (keypoints.cu)
    unsigned* nms_level;            // [slots] per-image level L: only pixels with score bits >= L are processed
    int* nms_flag;                  // [slots] 1 = redo this image from the plain threshold
    uint2* nms_list;                // [slots][GNB_NMS_LIST_CAP] (pixel index, score bits) of the pixels at or above the level
    int* nms_list_count;            // [slots] listed pixels, then [slots] "any image of the call starting at this slot to redo"
    unsigned* nms_sup;              // [slots][2][nms_sup_words] suppression bitmaps; a call uses the region of its first slot as [2][n][h][ceil(w / 32)]
    size_t nms_sup_words;           // words of ONE bitmap of ONE image at the configured maximum size (an odd aspect ratio that needs more uses the tile kernel)
    float* kp_xy;                   // [slots][K][2]
    float* kp_score;                // [slots][K]
    int* kp_count;                  // [slots]
    float* desc_f32;                // [slots][K][256]
    // matcher, per pair
    bf16* mproj;                    // [slots][K][256]
    float* mlogit;                  // [slots][K]   logsigmoid(z)
    float* row_lse;                 // [slots][K]   LSE over the other side
    float* best_val;                // [slots][K]
    int* best_idx;                  // [slots][K]
    int* match_idx;                 // [pairs][K][2]
    float* match_score;             // [pairs][K]
    int* match_count;               // [pairs]
    float* mkp_qry;                 // [pairs][K][2]
    float* mkp_ref;                 // [pairs][K][2]
    // PnP, per pair
    float* obj;                     // [pairs][K][3]
    float* hyp;                     // [pairs][iters][12]
    int* hyp_count;                 // [pairs][iters]
    uint8_t* inlier_mask;           // [pairs][K]
    int* range_flag;                // [pairs]
    double* kmat;                   // [pairs][9]
    double* affine;                 // [pairs][12]
    uint8_t* dem;                   // [pairs][h][w]
    PairOut* out_dev;               // [pairs]
    PairOut* out_host;              // pinned [pairs]
    // staging for the stage-level entry points
    float* stage_a;                 // generic device scratch (floats)
    size_t stage_a_floats;
    float* stage_b;
    size_t stage_b_floats;
    // reference-raster feature cache (pose_node.py:226-241 caches the raster's features per stamp)
    int cache_cap;
    long long* cache_ids;           // host [cache_cap], -1 = free
    unsigned long long* cache_lru;  // host [cache_cap]
    unsigned long long cache_clock;
    int cache_h, cache_w;           // raster geometry the cached features belong to
    float* c_kp_xy;                 // [cache_cap][K][2]
    int* c_kp_count;                // [cache_cap]
    bf16* c_mproj;                  // [cache_cap][K][256]
    float* c_mlogit;                // [cache_cap][K]
    float* c_desc;                  // [cache_cap][K][256] raw descriptors, allocated with the transformer layers (lightglue.cu)
    uint8_t* warp_buf;              // staging of gnb_rotate_crop's host-buffer path (warp.cu), grow-only
    size_t warp_bytes;
    void* lg_state;                 // LgState* (lightglue.cu): transformer layers in front of the head, NULL = head only
    void* tc_state;                 // TcState* (tc_common.cuh): tensor maps bound to this context's buffers
    int* tc_err_host;               // host-mapped word written by a timed-out tcgen05 pipeline wait (per context, portable)
    int* tc_err_dev;
    const void* attr_funcs[128];    // kernels whose dynamic shared-memory limit has been raised on THIS context's device
    int n_attr_funcs;
    float* match_w_f32;             // fp32-faithful mode: matcher head weights [256 out][256 in] / [256] in fp32
    float* match_mw_f32;
    float* mproj_f32;               // [slots][K][256] fp32 projected descriptors (fp32-faithful mode)
    bf16* mproj_x3;                 // [slots][K][hi: 256 | lo: 256] their split-bf16 form: operands of the tensor-core matcher
    float* col_pa;                  // K4 column partials [max_batch][ceil(K/128)][K]: pass 0 max
    float* col_pb;                  //                                                  pass 0 sum exp
    float* col_qa;                  //                                                  pass 1 best score
    float* col_qb;                  //                                                  pass 1 best row (int bits)
    float* c_mproj_f32;             // [cache_cap][K][256] their cached copies
    bf16* c_mproj_x3;               // [cache_cap][K][512]
    // profiling
    int prof_on;
    void* prof;  // ProfState*
};

#define GNB_SET_ERR(ctx, ...)                                   \
    do {                                                        \
        if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
    } while (0)

#define GNB_CUDA(ctx, expr)                                                                  \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            GNB_SET_ERR(ctx, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return GNB_E_CUDA;                                                               \
        }                                                                                    \
    } while (0)

#define GNB_LAUNCH_CHECK(ctx)                          \
    do {                                               \
        (ctx)->launches++;                             \
        GNB_CUDA(ctx, cudaGetLastError());             \
    } while (0)

// Per-kernel CUDA-event timing on the launching stream (bench.py roofline): when profiling is on,
// every GNB_KERNEL launch is bracketed by an event pair; gnb_profile_read aggregates by name.
void gnb_prof_begin(gnb_ctx* ctx, const char* name);
void gnb_prof_end(gnb_ctx* ctx);
#define GNB_KERNEL(ctx, name, ...)     \
    do {                               \
        gnb_prof_begin(ctx, name);     \
        __VA_ARGS__;                   \
        gnb_prof_end(ctx);             \
        GNB_LAUNCH_CHECK(ctx);         \
    } while (0)

// stream sync + check of the tcgen05 pipeline watchdog word
#define GNB_SYNC(ctx)                                              \
    do {                                                           \
        GNB_CUDA(ctx, cudaStreamSynchronize((ctx)->stream));       \
        int _rc = gnb_tc_err_check(ctx);                           \
        if (_rc) return _rc;                                       \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: remember per context (= per device) which
// kernels have it raised, so two contexts on two GPUs of one process both get it (api.cu).
cudaError_t gnb_func_smem_impl(gnb_ctx* ctx, const void* func, int bytes);
template <typename F>
static inline cudaError_t gnb_func_smem(gnb_ctx* ctx, F* func, int bytes) { return gnb_func_smem_impl(ctx, (const void*)func, bytes); }

// ---- stage functions implemented across the .cu files (all enqueue on ctx->stream) -----------
int gnb_conv_init(gnb_ctx* ctx, const float* blob_floats_dev);  // repack weights (device pointer to the blob's floats)
void gnb_conv_free(gnb_ctx* ctx);
// run the dense stack on cw.img (n images of h x w already resident)
int gnb_conv_forward(gnb_ctx* ctx, int n, int h, int w, int dense_desc);
// descriptors of the selected keypoints of `n` images (slots slot0..): on-demand tcgen05 head or dense+sample
int gnb_describe(gnb_ctx* ctx, int n, int h, int w, int slot0);
// tcgen05 implicit-GEMM conv (conv_tc.cu); returns GNB_E_INVALID if a layer shape is unsupported
int gnb_conv_tc_layer(gnb_ctx* ctx, const ConvLayer& L, const bf16* in, int n, int h, int w, bf16* out_bf,
                      float* out_f, int relu, int pool);
int gnb_conv_tc_init(gnb_ctx* ctx);

// keypoints.cu
int gnb_kp_select(gnb_ctx* ctx, const float* score, int n, int h, int w, int slot0);
int gnb_kp_sample(gnb_ctx* ctx, const float* dense, int n, int h, int w, int slot0);

// match.cu
int gnb_match_init(gnb_ctx* ctx, const float* proj_w, const float* proj_b, const float* m_w, const float* m_b);  // device pointers into the blob
void gnb_match_free(gnb_ctx* ctx);
// project descriptors of `n_slots` consecutive slots
int gnb_match_project(gnb_ctx* ctx, int slot0, int n_slots);
// match slot_a[p] against slot_b[p] for p in [0, pairs): fills match_idx/score/count + mkp_*
int gnb_match_pairs(gnb_ctx* ctx, int pairs, int slot_a0, int slot_b0, int stride_a = 1);
int gnb_match_tc_init(gnb_ctx* ctx);
void gnb_tc_state_free(gnb_ctx* ctx);
int gnb_knn_ratio(gnb_ctx* ctx, const float* dq, int nq, const float* dr, int nr, int dim, double ratio);
int gnb_match_tc_rowpass(gnb_ctx* ctx, int pairs, int slot_a0, int slot_b0, int stride_a, int pass);
int gnb_match_tc_pairpass(gnb_ctx* ctx, int pairs, int slot_a0, int slot_b0, int stride_a);

// pnp.cu
int gnb_pnp_pairs(gnb_ctx* ctx, int pairs, int dem_h, int dem_w, int has_dem, int ref_h, int ref_w, int do_tail,
                  int min_matches, int use_kp_counts, int stride_a = 1);
int gnb_tail_device(gnb_ctx* ctx, const double* r9, const double* t3, const double* affine12, int ref_h, int ref_w,
                    double* ecef3, double* quat4, double* lla3);
int gnb_ensure_stage(gnb_ctx* ctx, size_t floats_a, size_t floats_b);

// lightglue.cu
void gnb_lightglue_free(gnb_ctx* ctx);
// run the transformer layers in place on desc_f32 for pairs (slot_a0 + p, slot_b0 + p); no-op without layers
int gnb_lightglue_forward(gnb_ctx* ctx, int pairs, int slot_a0, int slot_b0, float ha, float wa, float hb, float wb);
