// api.cu — C ABI of libgisnav_b200.so (include/gisnav_b200.h): context lifetime, workspace, the
// three reference call sites (detectAndCompute / matcher / compute_pose), the fused batch path
// and the stage-isolated hooks used by the parity tests.
#include <utility>
#include "common.cuh"

#include <nvtx3/nvToolsExt.h>

#include <stdlib.h>

#include <new>
#include <vector>

// NVTX range around a stage of the path (K1 dense stack, K2/K3 keypoints, K4 matcher, K5/K6 PnP + tail): visible in
// Nsight Systems / ncu --nvtx; a no-op push/pop when no tool is attached
struct GnbRange {
    explicit GnbRange(const char* name) { nvtxRangePushA(name); }
    ~GnbRange() { nvtxRangePop(); }
};

static char g_create_err[512] = "";

// ---- per-kernel event timing ---------------------------------------------------------------------
struct ProfRec { const char* name; cudaEvent_t a, b; };
struct ProfState { std::vector<ProfRec> recs; std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool; };

void gnb_prof_begin(gnb_ctx* ctx, const char* name) {
    if (!ctx->prof_on) return;
    ProfState* ps = static_cast<ProfState*>(ctx->prof);
    ProfRec r;
    r.name = name;
    if (!ps->pool.empty()) { r.a = ps->pool.back().first; r.b = ps->pool.back().second; ps->pool.pop_back(); }
    else { cudaEventCreate(&r.a); cudaEventCreate(&r.b); }
    cudaEventRecord(r.a, ctx->stream);
    ps->recs.push_back(r);
}
void gnb_prof_end(gnb_ctx* ctx) {
    if (!ctx->prof_on) return;
    ProfState* ps = static_cast<ProfState*>(ctx->prof);
    cudaEventRecord(ps->recs.back().b, ctx->stream);
}

extern "C" int gnb_profile_enable(gnb_ctx* ctx, int on) {
    if (!ctx) return GNB_E_INVALID;
    if (!ctx->prof) ctx->prof = new ProfState();
    ctx->prof_on = on ? 1 : 0;
    return GNB_OK;
}

// Aggregate the recorded launches by kernel name.  names: cap entries of 64 chars.  Clears the log.
extern "C" int gnb_profile_read(gnb_ctx* ctx, char* names, float* total_ms, int64_t* launches, int cap, int* n_out) {
    if (!ctx || !names || !total_ms || !launches || !n_out) return GNB_E_INVALID;
    *n_out = 0;
    if (!ctx->prof) return GNB_OK;
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    GNB_SYNC(ctx);
    ProfState* ps = static_cast<ProfState*>(ctx->prof);
    int n = 0;
    for (auto& r : ps->recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        int j = 0;
        for (; j < n; ++j) if (strncmp(names + 64 * j, r.name, 63) == 0) break;
        if (j == n) {
            if (n >= cap) { ps->pool.push_back({r.a, r.b}); continue; }
            strncpy(names + 64 * j, r.name, 63); names[64 * j + 63] = 0; total_ms[j] = 0.f; launches[j] = 0; ++n;
        }
        total_ms[j] += ms; launches[j] += 1;
        ps->pool.push_back({r.a, r.b});
    }
    ps->recs.clear();
    *n_out = n;
    return GNB_OK;
}

cudaError_t gnb_func_smem_impl(gnb_ctx* ctx, const void* func, int bytes) {
    for (int i = 0; i < ctx->n_attr_funcs; ++i)
        if (ctx->attr_funcs[i] == func) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && ctx->n_attr_funcs < (int)(sizeof(ctx->attr_funcs) / sizeof(ctx->attr_funcs[0]))) ctx->attr_funcs[ctx->n_attr_funcs++] = func;
    return e;
}

extern "C" int gnb_default_config(gnb_config* cfg) {
    if (!cfg) return GNB_E_INVALID;
    memset(cfg, 0, sizeof(*cfg));
    cfg->max_keypoints = 1024;
    cfg->nms_radius = 4;
    cfg->keypoint_threshold = 0.005f;
    cfg->border = 4;
    cfg->match_threshold = 0.5f;
    cfg->min_matches = 15;
    cfg->ransac_iters = 2048;
    cfg->reproj_px = 8.0f;
    cfg->ransac_seed = 0;
    cfg->refine = 1;
    cfg->max_batch = 8;
    cfg->max_image_h = 1088;
    cfg->max_image_w = 1280;
    cfg->conv_impl = 0;   // tcgen05 implicit GEMM
    cfg->match_impl = 0;  // tcgen05 descriptor GEMM
    cfg->tile_cache = 32;
    cfg->precision = 0;   // bf16 operands (fast mode); 1 = split-bf16 operands, fp32-faithful
    return GNB_OK;
}

extern "C" const char* gnb_last_error(const gnb_ctx* ctx) { return ctx ? ctx->err : g_create_err; }
extern "C" int64_t gnb_launch_count(const gnb_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" void* gnb_stream(const gnb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int gnb_get_config(const gnb_ctx* ctx, gnb_config* out) {
    if (!ctx || !out) return GNB_E_INVALID;
    *out = ctx->cfg;
    return GNB_OK;
}

int gnb_ensure_stage(gnb_ctx* ctx, size_t fa, size_t fb) {
    if (fa > ctx->stage_a_floats) {
        if (ctx->stage_a) cudaFree(ctx->stage_a);
        ctx->stage_a = nullptr; ctx->stage_a_floats = 0;
        GNB_CUDA(ctx, cudaMalloc(&ctx->stage_a, fa * sizeof(float)));
        ctx->stage_a_floats = fa;
    }
    if (fb > ctx->stage_b_floats) {
        if (ctx->stage_b) cudaFree(ctx->stage_b);
        ctx->stage_b = nullptr; ctx->stage_b_floats = 0;
        GNB_CUDA(ctx, cudaMalloc(&ctx->stage_b, fb * sizeof(float)));
        ctx->stage_b_floats = fb;
    }
    return GNB_OK;
}

template <typename T>
static int dalloc(gnb_ctx* ctx, T** p, size_t count) {
    GNB_CUDA(ctx, cudaMalloc((void**)p, count * sizeof(T)));
    return GNB_OK;
}

#define GNB_OVERLAP_IMAGES 2

// the activation buffers of one pass over n images of at most px pixels
static int alloc_activations(gnb_ctx* ctx, ConvWorkspace& cw, size_t n, size_t px) {
    const size_t pf = ctx->cfg.precision == 1 ? 2 : 1;   // fp32-faithful mode: every activation is a (hi, lo) pair of bf16
    int rc = 0;
    rc |= dalloc(ctx, &cw.a1a, n * px * 64 * pf);
    rc |= dalloc(ctx, &cw.p1, n * px / 4 * 64 * pf);
    rc |= dalloc(ctx, &cw.a2a, n * px / 4 * 64 * pf);
    rc |= dalloc(ctx, &cw.p2, n * px / 16 * 64 * pf);
    rc |= dalloc(ctx, &cw.a3a, n * px / 16 * 128 * pf);
    rc |= dalloc(ctx, &cw.p3, n * px / 64 * 128 * pf);
    rc |= dalloc(ctx, &cw.a4a, n * px / 64 * 128 * pf);
    rc |= dalloc(ctx, &cw.a4b, n * px / 64 * 128 * pf);
    rc |= dalloc(ctx, &cw.apa, n * px / 64 * 256 * pf);
    rc |= dalloc(ctx, &cw.ada, n * px / 64 * 256 * pf);
    rc |= dalloc(ctx, &cw.semi, n * px / 64 * 65);
    rc |= dalloc(ctx, &cw.score, n * px);
    rc |= dalloc(ctx, &cw.dense, n * px / 64 * 256);
    return rc;
}

static int alloc_workspace(gnb_ctx* ctx) {
    const gnb_config& c = ctx->cfg;
    ConvWorkspace& cw = ctx->cw;
    const size_t n = c.max_batch, px = (size_t)c.max_image_h * c.max_image_w;
    cw.cap_images = c.max_batch;
    cw.cap_pixels = px;
    int rc = 0;
    rc |= dalloc(ctx, &cw.img_a, n * px);
    rc |= dalloc(ctx, &cw.img_b, n * px);
    cw.img = cw.img_a;
    rc |= alloc_activations(ctx, cw, n, px);
    // small batches (the per-message case of the reference node) run the raster chain on a second stream next to the frame
    // chain: that needs its own activation buffers, for up to GNB_OVERLAP_IMAGES images
    ctx->overlap_images = (int)(n < GNB_OVERLAP_IMAGES ? n : GNB_OVERLAP_IMAGES);
    ctx->cw2 = cw;
    rc |= alloc_activations(ctx, ctx->cw2, ctx->overlap_images, px);
    ctx->cw2.cap_images = ctx->overlap_images;
    if (rc) return GNB_E_CUDA;
    const size_t slots = 2 * n, k = c.max_keypoints, it = c.ransac_iters;
    ctx->kp_slots = (int)slots;
    rc |= dalloc(ctx, &ctx->cand_keys, slots * GNB_CAND_CAP);
    rc |= dalloc(ctx, &ctx->cand_count, slots);
    rc |= dalloc(ctx, &ctx->nms_hist, slots * 2048 + slots);   // + one ticket word per slot
    rc |= dalloc(ctx, &ctx->nms_level, slots);
    rc |= dalloc(ctx, &ctx->nms_flag, slots);
    rc |= dalloc(ctx, &ctx->nms_list, slots * GNB_NMS_LIST_CAP);
    rc |= dalloc(ctx, &ctx->nms_list_count, 2 * slots);   // counts, then the any-redo words
    ctx->nms_sup_words = (size_t)c.max_image_h * ((c.max_image_w + 31) / 32);   // one bitmap of one image
    rc |= dalloc(ctx, &ctx->nms_sup, slots * 2 * ctx->nms_sup_words);
    rc |= dalloc(ctx, &ctx->kp_xy, slots * k * 2);
    rc |= dalloc(ctx, &ctx->kp_score, slots * k);
    rc |= dalloc(ctx, &ctx->kp_count, slots);
    rc |= dalloc(ctx, &ctx->desc_f32, slots * k * 256);
    rc |= dalloc(ctx, &ctx->mproj, slots * k * 256);
    rc |= dalloc(ctx, &ctx->mlogit, slots * k);
    rc |= dalloc(ctx, &ctx->row_lse, slots * k);
    rc |= dalloc(ctx, &ctx->best_val, slots * k);
    rc |= dalloc(ctx, &ctx->best_idx, slots * k);
    rc |= dalloc(ctx, &ctx->match_idx, n * k * 2);
    rc |= dalloc(ctx, &ctx->match_score, n * k);
    rc |= dalloc(ctx, &ctx->match_count, n);
    rc |= dalloc(ctx, &ctx->mkp_qry, n * k * 2);
    rc |= dalloc(ctx, &ctx->mkp_ref, n * k * 2);
    rc |= dalloc(ctx, &ctx->obj, n * k * 3);
    rc |= dalloc(ctx, &ctx->hyp, n * it * 12);
    rc |= dalloc(ctx, &ctx->hyp_count, n * it);
    rc |= dalloc(ctx, &ctx->inlier_mask, n * k);
    rc |= dalloc(ctx, &ctx->range_flag, n);
    rc |= dalloc(ctx, &ctx->kmat, n * 9);
    rc |= dalloc(ctx, &ctx->affine, n * 12);
    rc |= dalloc(ctx, &ctx->dem, n * px);
    rc |= dalloc(ctx, &ctx->out_dev, n);
    ctx->cache_cap = c.tile_cache > (int)n ? c.tile_cache : (int)n;
    const size_t cc = ctx->cache_cap;
    rc |= dalloc(ctx, &ctx->c_kp_xy, cc * k * 2);
    rc |= dalloc(ctx, &ctx->c_kp_count, cc);
    rc |= dalloc(ctx, &ctx->c_mproj, cc * k * 256);
    rc |= dalloc(ctx, &ctx->c_mlogit, cc * k);
    if (c.precision == 1) {
        rc |= dalloc(ctx, &ctx->mproj_f32, slots * k * 256);
        rc |= dalloc(ctx, &ctx->c_mproj_f32, cc * k * 256);
        if (c.match_impl == 0) rc |= dalloc(ctx, &ctx->c_mproj_x3, cc * k * 512);
    }
    if (rc) return GNB_E_CUDA;
    if (c.precision == 1) GNB_CUDA(ctx, cudaMemset(ctx->mproj_f32, 0, slots * k * 256 * sizeof(float)));
    GNB_CUDA(ctx, cudaMemset(ctx->desc_f32, 0, slots * k * 256 * sizeof(float)));
    ctx->cache_ids = new long long[cc];
    ctx->cache_lru = new unsigned long long[cc];
    for (size_t i = 0; i < cc; ++i) { ctx->cache_ids[i] = -1; ctx->cache_lru[i] = 0; }
    ctx->cache_clock = 0;
    GNB_CUDA(ctx, cudaMemset(ctx->kp_count, 0, slots * sizeof(int)));
    GNB_CUDA(ctx, cudaMemset(ctx->nms_hist, 0, (slots * 2048 + slots) * sizeof(unsigned)));
    GNB_CUDA(ctx, cudaMemset(ctx->mproj, 0, slots * k * 256 * sizeof(bf16)));
    GNB_CUDA(ctx, cudaMemset(ctx->kp_xy, 0, slots * k * 2 * sizeof(float)));
    GNB_CUDA(ctx, cudaMallocHost((void**)&ctx->out_host, n * sizeof(PairOut)));
    return gnb_ensure_stage(ctx, 1 << 16, 1 << 16);
}

extern "C" void gnb_destroy(gnb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    gnb_lightglue_free(ctx);
    gnb_conv_free(ctx);
    gnb_match_free(ctx);
    ConvWorkspace& cw = ctx->cw;
    void* ptrs[] = {cw.img_a, cw.img_b, cw.a1a, cw.p1, cw.a2a, cw.p2, cw.a3a, cw.p3, cw.a4a, cw.a4b, cw.apa, cw.ada, cw.semi,
                    cw.score, cw.dense, ctx->cand_keys, ctx->cand_count, ctx->kp_xy, ctx->kp_score, ctx->kp_count,
                    ctx->desc_f32, ctx->mproj, ctx->mlogit, ctx->row_lse, ctx->best_val, ctx->best_idx, ctx->match_idx,
                    ctx->match_score, ctx->match_count, ctx->mkp_qry, ctx->mkp_ref, ctx->obj, ctx->hyp, ctx->hyp_count,
                    ctx->inlier_mask, ctx->range_flag, ctx->kmat, ctx->affine, ctx->dem, ctx->out_dev, ctx->stage_a,
                    ctx->stage_b, ctx->c_kp_xy, ctx->c_kp_count, ctx->c_mproj, ctx->c_mlogit, ctx->c_desc, ctx->warp_buf,
                    ctx->mproj_f32, ctx->c_mproj_f32, ctx->c_mproj_x3, ctx->nms_hist, ctx->nms_level, ctx->nms_flag, ctx->nms_list, ctx->nms_list_count, ctx->nms_sup};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    {
        ConvWorkspace& c2 = ctx->cw2;
        void* a[] = {c2.a1a, c2.p1, c2.a2a, c2.p2, c2.a3a, c2.p3, c2.a4a, c2.a4b, c2.apa, c2.ada, c2.semi, c2.score, c2.dense};
        void* b[] = {cw.a1a, cw.p1, cw.a2a, cw.p2, cw.a3a, cw.p3, cw.a4a, cw.a4b, cw.apa, cw.ada, cw.semi, cw.score, cw.dense};
        for (int i = 0; i < 13; ++i)
            if (a[i] && a[i] != b[i]) cudaFree(a[i]);
    }
    if (ctx->out_host) cudaFreeHost(ctx->out_host);
    gnb_tc_state_free(ctx);
    delete[] ctx->cache_ids;
    delete[] ctx->cache_lru;
    if (ctx->prof) {
        ProfState* ps = static_cast<ProfState*>(ctx->prof);
        for (auto& r : ps->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
        for (auto& p : ps->pool) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
        delete ps;
    }
    if (ctx->ev_frames) cudaEventDestroy(ctx->ev_frames);
    if (ctx->ev_tiles) cudaEventDestroy(ctx->ev_tiles);
    if (ctx->ev_params) cudaEventDestroy(ctx->ev_params);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int gnb_create(const gnb_config* cfg, const void* weights, size_t nbytes, int weights_on_device, int device,
                          gnb_ctx** out) {
    if (!cfg || !weights || !out) { snprintf(g_create_err, sizeof(g_create_err), "null argument"); return GNB_E_INVALID; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        snprintf(g_create_err, sizeof(g_create_err), "no CUDA device %d (found %d); this library has no CPU fallback", device, ndev);
        return GNB_E_NO_DEVICE;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10) {
        snprintf(g_create_err, sizeof(g_create_err), "device %d is sm_%d%d; libgisnav_b200 is built for sm_100a only", device,
                 prop.major, prop.minor);
        return GNB_E_NO_DEVICE;
    }
    if (cfg->max_keypoints < 16 || cfg->max_keypoints > GNB_MAX_KP || cfg->max_batch < 1 || cfg->ransac_iters < 1 ||
        cfg->max_image_h % 8 || cfg->max_image_w % 8 || cfg->max_image_h < 16 || cfg->max_image_w < 16) {
        snprintf(g_create_err, sizeof(g_create_err), "invalid config (K in [16,%d], sides multiple of 8)", GNB_MAX_KP);
        return GNB_E_INVALID;
    }
    if (cfg->precision != 0 && cfg->precision != 1) {
        snprintf(g_create_err, sizeof(g_create_err), "invalid config: precision must be 0 (bf16) or 1 (fp32-faithful)");
        return GNB_E_INVALID;
    }
    gnb_ctx* ctx = new (std::nothrow) gnb_ctx();
    if (!ctx) return GNB_E_INVALID;
    memset(ctx, 0, sizeof(*ctx));
    ctx->cfg = *cfg;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->err[0] = 0;
    int rc = GNB_OK;
    auto fail = [&](int code) {
        snprintf(g_create_err, sizeof(g_create_err), "%s", ctx->err);
        gnb_destroy(ctx);
        return code;
    };
    if (cudaSetDevice(device) != cudaSuccess) { GNB_SET_ERR(ctx, "cudaSetDevice failed"); return fail(GNB_E_CUDA); }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_frames, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_tiles, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_params, cudaEventDisableTiming) != cudaSuccess) {
        GNB_SET_ERR(ctx, "stream create failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(GNB_E_CUDA);
    }
    // weight blob: 16-byte header + floats (gisnav_b200/weights.py).  The floats are repacked ON THE DEVICE
    // (gnb_conv_init / gnb_match_init launch the repack kernels): a blob that arrived by NCCL broadcast is consumed
    // where it lies, a host blob is copied once.
    const size_t expect_floats = 1366914;
    if (nbytes != 16 + 4 * expect_floats) { GNB_SET_ERR(ctx, "bad weight blob size %zu (expected %zu)", nbytes, 16 + 4 * expect_floats); return fail(GNB_E_INVALID); }
    uint32_t hdr[4];
    if (cudaMemcpy(hdr, weights, 16, weights_on_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost) != cudaSuccess) {
        GNB_SET_ERR(ctx, "cannot read the weight blob header");
        return fail(GNB_E_CUDA);
    }
    if (memcmp(hdr, "GNBW", 4) != 0 || hdr[1] != 1 || hdr[2] != expect_floats) {
        GNB_SET_ERR(ctx, "bad weight blob (magic/version/size)");
        return fail(GNB_E_INVALID);
    }
    const float* fl = nullptr;   // device pointer to the blob's floats
    float* staged = nullptr;
    if (weights_on_device) {
        fl = reinterpret_cast<const float*>(static_cast<const uint8_t*>(weights) + 16);
    } else {
        if (cudaMalloc(&staged, 4 * expect_floats) != cudaSuccess ||
            cudaMemcpy(staged, static_cast<const uint8_t*>(weights) + 16, 4 * expect_floats, cudaMemcpyHostToDevice) != cudaSuccess) {
            GNB_SET_ERR(ctx, "cannot stage the weight blob on the device");
            if (staged) cudaFree(staged);
            return fail(GNB_E_CUDA);
        }
        fl = staged;
    }
    rc = gnb_conv_init(ctx, fl);
    const float* head = fl + (expect_floats - (256 * 256 + 256 + 256 + 1));
    if (!rc) rc = gnb_match_init(ctx, head, head + 256 * 256, head + 256 * 256 + 256, head + 256 * 256 + 512);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && !rc) { GNB_SET_ERR(ctx, "weight repack failed: %s", cudaGetErrorString(cudaGetLastError())); rc = GNB_E_CUDA; }
    if (staged) cudaFree(staged);
    if (rc) return fail(rc);
    if ((rc = alloc_workspace(ctx))) return fail(rc);
    if ((cfg->conv_impl == 0 || cfg->precision == 1) && (rc = gnb_conv_tc_init(ctx))) return fail(rc);
    if (cfg->match_impl == 0 && (rc = gnb_match_tc_init(ctx))) return fail(rc);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { GNB_SET_ERR(ctx, "init sync failed"); return fail(GNB_E_CUDA); }
    *out = ctx;
    return GNB_OK;
}

// ------------------------------------------------------------------------------------------------
static cudaMemcpyKind kind_in(int on_device) { return on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice; }
static cudaMemcpyKind kind_out(int on_device) { return on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost; }

static int check_image(gnb_ctx* ctx, int h, int w) {
    if (h <= 0 || w <= 0 || (h % 8) || (w % 8)) {
        GNB_SET_ERR(ctx, "image sides must be positive multiples of 8 (got %dx%d)", h, w);
        return GNB_E_INVALID;
    }
    if ((size_t)h * w > ctx->cw.cap_pixels) {
        GNB_SET_ERR(ctx, "image %dx%d exceeds the workspace (%d x %d)", h, w, ctx->cfg.max_image_h, ctx->cfg.max_image_w);
        return GNB_E_CAPACITY;
    }
    return GNB_OK;
}

static int read_count(gnb_ctx* ctx, const int* dptr, int* out) {
    GNB_CUDA(ctx, cudaMemcpyAsync(out, dptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GNB_SYNC(ctx);
    return GNB_OK;
}

extern "C" int gnb_extract(gnb_ctx* ctx, const uint8_t* image, int h, int w, int stride, int on_device, float* out_xy,
                           float* out_score, float* out_desc, int cap, int* n_out) {
    if (!ctx || !image || !n_out) return GNB_E_INVALID;
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = check_image(ctx, h, w))) return rc;
    if (stride < w) { GNB_SET_ERR(ctx, "stride < width"); return GNB_E_INVALID; }
    GNB_CUDA(ctx, cudaMemcpy2DAsync(ctx->cw.img, w, image, stride, w, h, kind_in(on_device), ctx->stream));
    if ((rc = gnb_conv_forward(ctx, 1, h, w, 0))) return rc;
    if ((rc = gnb_kp_select(ctx, ctx->cw.score, 1, h, w, 0))) return rc;
    if ((rc = gnb_describe(ctx, 1, h, w, 0))) return rc;
    int n = 0;
    if ((rc = read_count(ctx, ctx->kp_count, &n))) return rc;
    if (n < 0) { GNB_SET_ERR(ctx, "NMS candidate buffer overflow"); return GNB_E_CAPACITY; }
    n = n < cap ? n : cap;
    *n_out = n;
    if (n > 0) {
        if (out_xy) GNB_CUDA(ctx, cudaMemcpyAsync(out_xy, ctx->kp_xy, sizeof(float) * 2 * n, kind_out(on_device), ctx->stream));
        if (out_score) GNB_CUDA(ctx, cudaMemcpyAsync(out_score, ctx->kp_score, sizeof(float) * n, kind_out(on_device), ctx->stream));
        if (out_desc) GNB_CUDA(ctx, cudaMemcpyAsync(out_desc, ctx->desc_f32, sizeof(float) * 256 * n, kind_out(on_device), ctx->stream));
        GNB_SYNC(ctx);
    }
    return GNB_OK;
}

__global__ void widen_idx_kernel(const int* __restrict__ in, long long* __restrict__ out, int n2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n2) out[i] = (long long)in[i];
}

struct KpArgs { const float *kp_a, *kp_b; float ha, wa, hb, wb; };

static int load_descs(gnb_ctx* ctx, const float* desc_a, int n_a, const float* desc_b, int n_b, int on_device,
                      const KpArgs* kps = nullptr) {
    const int k = ctx->cfg.max_keypoints, sb = ctx->cfg.max_batch;
    if (ctx->lg_state && !kps) {
        GNB_SET_ERR(ctx, "the matcher has transformer layers: keypoints are required (gnb_match_lightglue)");
        return GNB_E_INVALID;
    }
    if (n_a < 0 || n_b < 0 || n_a > k || n_b > k) {
        GNB_SET_ERR(ctx, "descriptor count exceeds max_keypoints=%d", k);
        return GNB_E_CAPACITY;
    }
    if (n_a) GNB_CUDA(ctx, cudaMemcpyAsync(ctx->desc_f32, desc_a, sizeof(float) * 256 * n_a, kind_in(on_device), ctx->stream));
    if (n_b) GNB_CUDA(ctx, cudaMemcpyAsync(ctx->desc_f32 + (size_t)sb * k * 256, desc_b, sizeof(float) * 256 * n_b, kind_in(on_device), ctx->stream));
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->kp_count, &n_a, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->kp_count + sb, &n_b, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    if (kps && n_a && n_b) {
        GNB_CUDA(ctx, cudaMemcpyAsync(ctx->kp_xy, kps->kp_a, sizeof(float) * 2 * n_a, kind_in(on_device), ctx->stream));
        GNB_CUDA(ctx, cudaMemcpyAsync(ctx->kp_xy + (size_t)sb * k * 2, kps->kp_b, sizeof(float) * 2 * n_b, kind_in(on_device), ctx->stream));
    }
    GNB_SYNC(ctx);  // n_a/n_b live on the caller's stack
    int rc;
    if (kps && (rc = gnb_lightglue_forward(ctx, 1, 0, sb, kps->ha, kps->wa, kps->hb, kps->wb))) return rc;
    if ((rc = gnb_match_project(ctx, 0, 1))) return rc;
    if ((rc = gnb_match_project(ctx, sb, 1))) return rc;
    return GNB_OK;
}

static int match_impl(gnb_ctx* ctx, const float* desc_a, int n_a, const float* desc_b, int n_b, int on_device, const KpArgs* kps,
                      int64_t* out_idx, float* out_score, int cap, int* n_out);

extern "C" int gnb_match(gnb_ctx* ctx, const float* desc_a, int n_a, const float* desc_b, int n_b, int on_device,
                         int64_t* out_idx, float* out_score, int cap, int* n_out) {
    return match_impl(ctx, desc_a, n_a, desc_b, n_b, on_device, nullptr, out_idx, out_score, cap, n_out);
}

extern "C" int gnb_match_lightglue(gnb_ctx* ctx, const float* desc_a, const float* kp_a, int n_a, float h_a, float w_a,
                                   const float* desc_b, const float* kp_b, int n_b, float h_b, float w_b, int on_device,
                                   int64_t* out_idx, float* out_score, int cap, int* n_out) {
    if ((n_a > 0 && !kp_a) || (n_b > 0 && !kp_b)) return GNB_E_INVALID;
    KpArgs kps{kp_a, kp_b, h_a, w_a, h_b, w_b};
    return match_impl(ctx, desc_a, n_a, desc_b, n_b, on_device, &kps, out_idx, out_score, cap, n_out);
}

// descriptors of the last gnb_match / gnb_match_lightglue call after the transformer layers (parity hook)
extern "C" int gnb_refined_descriptors(gnb_ctx* ctx, int side, float* out, int n) {
    if (!ctx || !out || n < 0 || n > ctx->cfg.max_keypoints || side < 0 || side > 1) return GNB_E_INVALID;
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t off = side ? (size_t)ctx->cfg.max_batch * ctx->cfg.max_keypoints * 256 : 0;
    if (n) GNB_CUDA(ctx, cudaMemcpyAsync(out, ctx->desc_f32 + off, sizeof(float) * 256 * n, cudaMemcpyDeviceToHost, ctx->stream));
    GNB_SYNC(ctx);
    return GNB_OK;
}

static int match_impl(gnb_ctx* ctx, const float* desc_a, int n_a, const float* desc_b, int n_b, int on_device, const KpArgs* kps,
                      int64_t* out_idx, float* out_score, int cap, int* n_out) {
    if (!ctx || !n_out) return GNB_E_INVALID;
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    *n_out = 0;
    if (n_a == 0 || n_b == 0) return GNB_OK;
    if (!desc_a || !desc_b) return GNB_E_INVALID;
    int rc;
    if ((rc = load_descs(ctx, desc_a, n_a, desc_b, n_b, on_device, kps))) return rc;
    if ((rc = gnb_match_pairs(ctx, 1, 0, ctx->cfg.max_batch))) return rc;
    int n = 0;
    if ((rc = read_count(ctx, ctx->match_count, &n))) return rc;
    n = n < cap ? n : cap;
    *n_out = n;
    if (n > 0) {
        if (out_idx) {
            if (on_device) {
                GNB_KERNEL(ctx, "widen_idx_kernel", widen_idx_kernel<<<ceil_div(2 * n, 256), 256, 0, ctx->stream>>>(ctx->match_idx, (long long*)out_idx, 2 * n));
            } else {
                std::vector<int> tmp(2 * n);
                GNB_CUDA(ctx, cudaMemcpyAsync(tmp.data(), ctx->match_idx, sizeof(int) * 2 * n, cudaMemcpyDeviceToHost, ctx->stream));
                GNB_SYNC(ctx);
                for (int i = 0; i < 2 * n; ++i) out_idx[i] = tmp[i];
            }
        }
        if (out_score) GNB_CUDA(ctx, cudaMemcpyAsync(out_score, ctx->match_score, sizeof(float) * n, kind_out(on_device), ctx->stream));
        GNB_SYNC(ctx);
    }
    return GNB_OK;
}

// TwistNode's matcher: brute-force 2-NN + ratio test (twist_node.py:248,263-267)
extern "C" int gnb_knn_ratio_match(gnb_ctx* ctx, const float* desc_q, int n_q, const float* desc_r, int n_r, int dim, double ratio,
                                   int on_device, int64_t* out_idx, float* out_dist, int cap, int* n_out) {
    if (!ctx || !n_out) return GNB_E_INVALID;
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    *n_out = 0;
    if (n_q == 0 || n_r < 2) return GNB_OK;   // knnMatch(k=2) yields no (m, n) pairs
    if (!desc_q || !desc_r || dim < 1 || dim > 256) return GNB_E_INVALID;
    const int k = ctx->cfg.max_keypoints;
    if (n_q > k || n_r > k) { GNB_SET_ERR(ctx, "descriptor count exceeds max_keypoints=%d", k); return GNB_E_CAPACITY; }
    if (ctx->cfg.match_impl != 0) { GNB_SET_ERR(ctx, "gnb_knn_ratio_match needs the tcgen05 matcher (match_impl = 0)"); return GNB_E_INVALID; }
    GnbRange range("gnb_knn_ratio_match");
    int rc;
    const float *dq = desc_q, *dr = desc_r;
    if (!on_device) {
        if ((rc = gnb_ensure_stage(ctx, (size_t)n_q * dim, (size_t)n_r * dim))) return rc;
        GNB_CUDA(ctx, cudaMemcpyAsync(ctx->stage_a, desc_q, sizeof(float) * n_q * dim, cudaMemcpyHostToDevice, ctx->stream));
        GNB_CUDA(ctx, cudaMemcpyAsync(ctx->stage_b, desc_r, sizeof(float) * n_r * dim, cudaMemcpyHostToDevice, ctx->stream));
        dq = ctx->stage_a; dr = ctx->stage_b;
    }
    const int sb = ctx->cfg.max_batch;
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->kp_count, &n_q, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->kp_count + sb, &n_r, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    GNB_SYNC(ctx);
    if ((rc = gnb_knn_ratio(ctx, dq, n_q, dr, n_r, dim, ratio))) return rc;
    int n = 0;
    if ((rc = read_count(ctx, ctx->match_count, &n))) return rc;
    n = n < cap ? n : cap;
    *n_out = n;
    if (n > 0) {
        if (on_device) {
            if (out_idx) GNB_KERNEL(ctx, "widen_idx_kernel", widen_idx_kernel<<<ceil_div(2 * n, 256), 256, 0, ctx->stream>>>(ctx->match_idx, (long long*)out_idx, 2 * n));
            if (out_dist) GNB_CUDA(ctx, cudaMemcpyAsync(out_dist, ctx->match_score, sizeof(float) * n, cudaMemcpyDeviceToDevice, ctx->stream));
            GNB_SYNC(ctx);
        } else {
            std::vector<int> tmp(2 * n);
            GNB_CUDA(ctx, cudaMemcpyAsync(tmp.data(), ctx->match_idx, sizeof(int) * 2 * n, cudaMemcpyDeviceToHost, ctx->stream));
            if (out_dist) GNB_CUDA(ctx, cudaMemcpyAsync(out_dist, ctx->match_score, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
            GNB_SYNC(ctx);
            if (out_idx) for (int i = 0; i < 2 * n; ++i) out_idx[i] = tmp[i];
        }
    }
    return GNB_OK;
}

static void pairout_to_result(const PairOut& p, gnb_pose_result* r) {
    r->status = p.status; r->n_kp_qry = p.n_kp_qry; r->n_kp_ref = p.n_kp_ref; r->n_matches = p.n_matches;
    r->n_inliers = p.n_inliers; r->best_hypothesis = p.best_hypothesis;
    memcpy(r->r, p.r, sizeof(p.r)); memcpy(r->t, p.t, sizeof(p.t)); memcpy(r->ecef, p.ecef, sizeof(p.ecef));
    memcpy(r->quat, p.quat, sizeof(p.quat)); memcpy(r->lla, p.lla, sizeof(p.lla));
}

extern "C" int gnb_solve_pnp(gnb_ctx* ctx, const float* mkp_qry, const float* mkp_ref, int n, const uint8_t* dem,
                             int dem_h, int dem_w, const double* k9, int on_device, double* out_r9, double* out_t3,
                             uint8_t* out_inlier_mask, int* n_inliers) {
    if (!ctx || !k9 || n < 0 || (n > 0 && (!mkp_qry || !mkp_ref))) return GNB_E_INVALID;
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n > ctx->cfg.max_keypoints) { GNB_SET_ERR(ctx, "n=%d exceeds max_keypoints", n); return GNB_E_CAPACITY; }
    if (dem && (size_t)dem_h * dem_w > ctx->cw.cap_pixels) { GNB_SET_ERR(ctx, "DEM exceeds the workspace"); return GNB_E_CAPACITY; }
    if (n_inliers) *n_inliers = 0;
    if (n < 4) return GNB_SOFT_PNP_FAILED;  // cv2.solvePnPRansac returns False below the minimal set
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->mkp_qry, mkp_qry, sizeof(float) * 2 * n, kind_in(on_device), ctx->stream));
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->mkp_ref, mkp_ref, sizeof(float) * 2 * n, kind_in(on_device), ctx->stream));
    if (dem) GNB_CUDA(ctx, cudaMemcpyAsync(ctx->dem, dem, (size_t)dem_h * dem_w, kind_in(on_device), ctx->stream));
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->kmat, k9, sizeof(double) * 9, kind_in(on_device), ctx->stream));
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->match_count, &n, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    GNB_SYNC(ctx);
    int rc;
    if ((rc = gnb_pnp_pairs(ctx, 1, dem_h, dem_w, dem ? 1 : 0, 0, 0, 0, 4, 0))) return rc;
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->out_host, ctx->out_dev, sizeof(PairOut), cudaMemcpyDeviceToHost, ctx->stream));
    GNB_SYNC(ctx);
    const PairOut& p = ctx->out_host[0];
    if (p.status == GNB_E_RANGE) { GNB_SET_ERR(ctx, "reference keypoint outside the DEM raster"); return GNB_E_RANGE; }
    if (n_inliers) *n_inliers = p.n_inliers;
    if (out_r9) memcpy(out_r9, p.r, sizeof(p.r));
    if (out_t3) memcpy(out_t3, p.t, sizeof(p.t));
    if (out_inlier_mask) {
        GNB_CUDA(ctx, cudaMemcpyAsync(out_inlier_mask, ctx->inlier_mask, n, kind_out(on_device), ctx->stream));
        GNB_SYNC(ctx);
    }
    return p.status;
}

extern "C" int gnb_geodetic_tail(gnb_ctx* ctx, const double* r9, const double* t3, const double* affine12, int ref_h,
                                 int ref_w, double* out_ecef3, double* out_quat4, double* out_lla3) {
    if (!ctx || !r9 || !t3 || !affine12 || !out_ecef3 || !out_quat4 || !out_lla3) return GNB_E_INVALID;
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    return gnb_tail_device(ctx, r9, t3, affine12, ref_h, ref_w, out_ecef3, out_quat4, out_lla3);
}

static int pose_batch_impl(gnb_ctx* ctx, int batch, const uint8_t* frames, int hq, int wq, const uint8_t* tiles, int ht, int wt,
                           const uint8_t* dems, const double* k9, const double* affine12, int on_device, gnb_pose_result* results);

extern "C" int gnb_pose_batch(gnb_ctx* ctx, int batch, const uint8_t* frames, int hq, int wq, const uint8_t* tiles, int ht,
                              int wt, const uint8_t* dems, const double* k9, const double* affine12, int on_device,
                              gnb_pose_result* results) {
    if (!ctx || !frames || !tiles || !k9 || !affine12 || !results || batch < 1) return GNB_E_INVALID;
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    GnbRange range("gnb_pose_batch");
    const int rc = pose_batch_impl(ctx, batch, frames, hq, wq, tiles, ht, wt, dems, k9, affine12, on_device, results);
    if (rc < 0) {
        // an error path must not leave copies of the caller's buffers in flight: the caller may free or reuse them
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamSynchronize(ctx->stream2);
        cudaStreamSynchronize(ctx->stream);
    }
    return rc;
}

static int pose_batch_impl(gnb_ctx* ctx, int batch, const uint8_t* frames, int hq, int wq, const uint8_t* tiles, int ht, int wt,
                           const uint8_t* dems, const double* k9, const double* affine12, int on_device, gnb_pose_result* results) {
    if (batch > ctx->cfg.max_batch) { GNB_SET_ERR(ctx, "batch %d exceeds max_batch %d", batch, ctx->cfg.max_batch); return GNB_E_CAPACITY; }
    int rc;
    if ((rc = check_image(ctx, hq, wq)) || (rc = check_image(ctx, ht, wt))) return rc;
    const int sb = ctx->cfg.max_batch;
    const cudaMemcpyKind kin = kind_in(on_device);
    // staging: all copies go to the copy stream; the compute stream waits on events, so the raster /
    // DEM / parameter copies overlap with the conv pass over the query frames
    cudaStream_t cs = ctx->copy_stream;
    ConvWorkspace& cw = ctx->cw;
    GNB_CUDA(ctx, cudaMemcpyAsync(cw.img_a, frames, (size_t)batch * hq * wq, kin, cs));
    GNB_CUDA(ctx, cudaEventRecord(ctx->ev_frames, cs));
    GNB_CUDA(ctx, cudaMemcpyAsync(cw.img_b, tiles, (size_t)batch * ht * wt, kin, cs));
    GNB_CUDA(ctx, cudaEventRecord(ctx->ev_tiles, cs));
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->kmat, k9, sizeof(double) * 9 * batch, kin, cs));
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->affine, affine12, sizeof(double) * 12 * batch, kin, cs));
    if (dems) GNB_CUDA(ctx, cudaMemcpyAsync(ctx->dem, dems, (size_t)batch * ht * wt, kin, cs));
    GNB_CUDA(ctx, cudaEventRecord(ctx->ev_params, cs));
    // Small batches (one pair per message is what the reference node sends): most kernels of one image leave SMs idle
    // (topk: one CTA; the NMS passes, the descriptor head, the 1/8-resolution layers: a fraction of a wave), so the raster
    // chain K1-K3 runs on a second stream with its own activation buffers, next to the frame chain, and joins before K4.
    // Large batches fill the GPU by themselves and stay serial: co-running kernels there was measured slower.
    static const int no_overlap = getenv("GNB_NO_OVERLAP") ? atoi(getenv("GNB_NO_OVERLAP")) : 0;
    const bool overlap = batch <= ctx->overlap_images && !ctx->prof_on && !no_overlap;
    struct Restore {   // the raster chain borrows ctx->stream / ctx->cw: put them back on every exit path
        gnb_ctx* c; cudaStream_t s; bool active;
        ~Restore() { if (active) { c->stream = s; std::swap(c->cw, c->cw2); } }
    } restore{ctx, ctx->stream, false};
    if (overlap) {   // whatever an earlier call left running on the main stream is ordered before the second stream's work
        GNB_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->stream));
        GNB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_join, 0));
    }
    // query frames -> slots [0, batch)
    GNB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_frames, 0));
    cw.img = cw.img_a;
    {
        GnbRange r1("K1 dense stack (frames)");
        if ((rc = gnb_conv_forward(ctx, batch, hq, wq, 0))) return rc;
    }
    {
        GnbRange r2("K2+K3 keypoints + descriptors (frames)");
        if ((rc = gnb_kp_select(ctx, ctx->cw.score, batch, hq, wq, 0))) return rc;
        if ((rc = gnb_describe(ctx, batch, hq, wq, 0))) return rc;
    }
    // reference rasters -> slots [max_batch, max_batch + batch)
    if (overlap) {
        std::swap(ctx->cw, ctx->cw2);
        ctx->stream = ctx->stream2;
        restore.active = true;
    }
    GNB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_tiles, 0));
    {
        GnbRange r1("K1 dense stack (rasters)");
        cw.img = cw.img_b;
        rc = gnb_conv_forward(ctx, batch, ht, wt, 0);
        cw.img = cw.img_a;
        if (rc) return rc;
    }
    {
        GnbRange r2("K2+K3 keypoints + descriptors (rasters)");
        if ((rc = gnb_kp_select(ctx, ctx->cw.score, batch, ht, wt, sb))) return rc;
        if ((rc = gnb_describe(ctx, batch, ht, wt, sb))) return rc;
    }
    if (overlap) {
        GNB_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->stream2));
        ctx->stream = restore.s;
        std::swap(ctx->cw, ctx->cw2);
        restore.active = false;
        GNB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    }
    GNB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_params, 0));
    {
        GnbRange r4("K4 matcher");
        // transformer layers of the reference matcher (pose_node.py:109-121), when a layer blob is loaded
        if ((rc = gnb_lightglue_forward(ctx, batch, 0, sb, (float)hq, (float)wq, (float)ht, (float)wt))) return rc;
        if ((rc = gnb_match_project(ctx, 0, batch))) return rc;
        if ((rc = gnb_match_project(ctx, sb, batch))) return rc;
        if ((rc = gnb_match_pairs(ctx, batch, 0, sb))) return rc;
    }
    GnbRange r5("K5+K6 PnP/RANSAC + refit + WGS84 tail");
    if ((rc = gnb_pnp_pairs(ctx, batch, ht, wt, dems ? 1 : 0, ht, wt, 1, ctx->cfg.min_matches, 1))) return rc;
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->out_host, ctx->out_dev, sizeof(PairOut) * batch, cudaMemcpyDeviceToHost, ctx->stream));
    GNB_SYNC(ctx);
    for (int b = 0; b < batch; ++b) {
        pairout_to_result(ctx->out_host[b], &results[b]);
        if (results[b].n_kp_qry < 0 || results[b].n_kp_ref < 0) {
            GNB_SET_ERR(ctx, "NMS candidate buffer overflow in pair %d", b);
            return GNB_E_CAPACITY;
        }
        if (results[b].status == GNB_E_RANGE) { GNB_SET_ERR(ctx, "reference keypoint outside the DEM in pair %d", b); return GNB_E_RANGE; }
    }
    return GNB_OK;
}

// ------------------------------------------------------------------------------------------------
// Wire-format ingest: the query side of an OrthoStereoImage message arrives as PRE-EXTRACTED keypoints, one packed record
// per keypoint (x, y, z, size, angle f32 + descriptor f32[D]; KEYPOINT_DTYPE, ros/gisnav/gisnav/core/_shared.py:26-35,
// decoded with np.frombuffer at pose_node.py:207-213).  The record bytes are copied to the device as they are and
// unpacked there into keypoint slot 0; the reference raster is extracted on the device as usual.
__global__ void __launch_bounds__(256) unpack_records_kernel(const float* __restrict__ rec, int n, int step_floats, int desc_dim,
                                                             float* __restrict__ kp_xy, float* __restrict__ desc) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= n) return;
    const float* p = rec + (size_t)r * step_floats;
    if (lane < 2) kp_xy[r * 2 + lane] = p[lane];
    for (int c = lane; c < 256; c += 32) desc[(size_t)r * 256 + c] = c < desc_dim ? p[5 + c] : 0.f;
}

extern "C" int gnb_pose_from_records(gnb_ctx* ctx, const void* records, int n_records, int point_step, int desc_dim, int hq, int wq,
                                     const uint8_t* reference, int ht, int wt, const uint8_t* dem, const double* k9,
                                     const double* affine12, gnb_pose_result* result) {
    if (!ctx || !reference || !k9 || !affine12 || !result || n_records < 0 || (n_records > 0 && !records)) return GNB_E_INVALID;
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (desc_dim != GNB_DESC_DIM || point_step != 4 * (5 + desc_dim)) {
        GNB_SET_ERR(ctx, "records must carry %d-d descriptors with point_step %d (got D=%d, step %d): 128-d SIFT records cannot feed this matcher head",
                    GNB_DESC_DIM, 4 * (5 + GNB_DESC_DIM), desc_dim, point_step);
        return GNB_E_INVALID;
    }
    const int k = ctx->cfg.max_keypoints, sb = ctx->cfg.max_batch;
    if (n_records > k) { GNB_SET_ERR(ctx, "%d records exceed max_keypoints=%d", n_records, k); return GNB_E_CAPACITY; }
    int rc;
    if ((rc = check_image(ctx, ht, wt))) return rc;
    GnbRange range("gnb_pose_from_records");
    ConvWorkspace& cw = ctx->cw;
    const size_t rec_floats = (size_t)n_records * (point_step / 4);
    if ((rc = gnb_ensure_stage(ctx, rec_floats ? rec_floats : 1, 0))) return rc;
    if (n_records) GNB_CUDA(ctx, cudaMemcpyAsync(ctx->stage_a, records, rec_floats * 4, cudaMemcpyHostToDevice, ctx->stream));
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->kp_count, &n_records, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    GNB_CUDA(ctx, cudaMemcpyAsync(cw.img_b, reference, (size_t)ht * wt, cudaMemcpyHostToDevice, ctx->stream));
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->kmat, k9, sizeof(double) * 9, cudaMemcpyHostToDevice, ctx->stream));
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->affine, affine12, sizeof(double) * 12, cudaMemcpyHostToDevice, ctx->stream));
    if (dem) GNB_CUDA(ctx, cudaMemcpyAsync(ctx->dem, dem, (size_t)ht * wt, cudaMemcpyHostToDevice, ctx->stream));
    GNB_SYNC(ctx);   // n_records lives on this stack frame
    if (n_records)
        GNB_KERNEL(ctx, "unpack_records_kernel", unpack_records_kernel<<<ceil_div(n_records * 32, 256), 256, 0, ctx->stream>>>(
            ctx->stage_a, n_records, point_step / 4, desc_dim, ctx->kp_xy, ctx->desc_f32));
    cw.img = cw.img_b;
    rc = gnb_conv_forward(ctx, 1, ht, wt, 0);
    cw.img = cw.img_a;
    if (rc) return rc;
    if ((rc = gnb_kp_select(ctx, cw.score, 1, ht, wt, sb))) return rc;
    if ((rc = gnb_describe(ctx, 1, ht, wt, sb))) return rc;
    if ((rc = gnb_lightglue_forward(ctx, 1, 0, sb, (float)(hq > 0 ? hq : ht), (float)(wq > 0 ? wq : wt), (float)ht, (float)wt))) return rc;
    if ((rc = gnb_match_project(ctx, 0, 1))) return rc;
    if ((rc = gnb_match_project(ctx, sb, 1))) return rc;
    if ((rc = gnb_match_pairs(ctx, 1, 0, sb))) return rc;
    if ((rc = gnb_pnp_pairs(ctx, 1, ht, wt, dem ? 1 : 0, ht, wt, 1, ctx->cfg.min_matches, 1))) return rc;
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->out_host, ctx->out_dev, sizeof(PairOut), cudaMemcpyDeviceToHost, ctx->stream));
    GNB_SYNC(ctx);
    pairout_to_result(ctx->out_host[0], result);
    if (result->n_kp_ref < 0) { GNB_SET_ERR(ctx, "NMS candidate buffer overflow"); return GNB_E_CAPACITY; }
    if (result->status == GNB_E_RANGE) { GNB_SET_ERR(ctx, "reference keypoint outside the DEM"); return GNB_E_RANGE; }
    return GNB_OK;
}

// ------------------------------------------------------------------------------------------------
// candidate search with a reference-raster feature cache
extern "C" int gnb_cache_clear(gnb_ctx* ctx) {
    if (!ctx) return GNB_E_INVALID;
    for (int i = 0; i < ctx->cache_cap; ++i) { ctx->cache_ids[i] = -1; ctx->cache_lru[i] = 0; }
    return GNB_OK;
}

static void cache_copy(gnb_ctx* ctx, int entry, int slot, bool to_cache) {
    const size_t k = ctx->cfg.max_keypoints;
    auto cp = [&](void* cache_p, void* slot_p, size_t bytes) {
        cudaMemcpyAsync(to_cache ? cache_p : slot_p, to_cache ? slot_p : cache_p, bytes, cudaMemcpyDeviceToDevice, ctx->stream);
    };
    cp(ctx->c_kp_xy + (size_t)entry * k * 2, ctx->kp_xy + (size_t)slot * k * 2, k * 2 * sizeof(float));
    cp(ctx->c_kp_count + entry, ctx->kp_count + slot, sizeof(int));
    if (ctx->lg_state) {
        // with transformer layers the cacheable part of a raster is what the extractor produced (keypoints + raw
        // descriptors): the refined features depend on the query it is paired with
        cp(ctx->c_desc + (size_t)entry * k * 256, ctx->desc_f32 + (size_t)slot * k * 256, k * 256 * sizeof(float));
        return;
    }
    if (ctx->cfg.precision == 1) {
        cp(ctx->c_mproj_f32 + (size_t)entry * k * 256, ctx->mproj_f32 + (size_t)slot * k * 256, k * 256 * sizeof(float));
        if (ctx->mproj_x3) cp(ctx->c_mproj_x3 + (size_t)entry * k * 512, ctx->mproj_x3 + (size_t)slot * k * 512, k * 512 * sizeof(bf16));
    } else
        cp(ctx->c_mproj + (size_t)entry * k * 256, ctx->mproj + (size_t)slot * k * 256, k * 256 * sizeof(bf16));
    cp(ctx->c_mlogit + (size_t)entry * k, ctx->mlogit + (size_t)slot * k, k * sizeof(float));
}

// cache -> slots for all candidates of a call in ONE launch (the per-array cudaMemcpyAsync form above costs ~40 launches
// per frame at 8 candidates).  Entry indices travel in the kernel parameters.
#define GNB_GATHER_MAX 32
struct CacheGather {
    int n, sb;
    int entry[GNB_GATHER_MAX];
    int n_arr;
    const char* c[5];      // cache arrays, [entry][bytes]
    char* s[5];            // slot arrays,  [slot][bytes]
    size_t bytes[5];
};

__global__ void __launch_bounds__(256) cache_gather_kernel(const CacheGather g) {
    const int i = blockIdx.y;
    for (int a = 0; a < g.n_arr; ++a) {
        const size_t nb = g.bytes[a];
        const char* src = g.c[a] + (size_t)g.entry[i] * nb;
        char* dst = g.s[a] + (size_t)(g.sb + i) * nb;
        if ((nb & 15) == 0) {
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* d4 = reinterpret_cast<uint4*>(dst);
            for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < nb / 16; j += (size_t)gridDim.x * blockDim.x) d4[j] = s4[j];
        } else {   // every array here is made of 4-byte words
            const uint32_t* s1 = reinterpret_cast<const uint32_t*>(src);
            uint32_t* d1 = reinterpret_cast<uint32_t*>(dst);
            for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < nb / 4; j += (size_t)gridDim.x * blockDim.x) d1[j] = s1[j];
        }
    }
}

static int cache_gather(gnb_ctx* ctx, const int* entry, int n, int sb) {
    const size_t k = ctx->cfg.max_keypoints;
    for (int i0 = 0; i0 < n; i0 += GNB_GATHER_MAX) {
        CacheGather g;
        g.n = n - i0 < GNB_GATHER_MAX ? n - i0 : GNB_GATHER_MAX;
        g.sb = sb + i0;
        for (int i = 0; i < g.n; ++i) g.entry[i] = entry[i0 + i];
        int a = 0;
        auto add = [&](const void* c, void* s_, size_t bytes) { g.c[a] = (const char*)c; g.s[a] = (char*)s_; g.bytes[a] = bytes; ++a; };
        add(ctx->c_kp_xy, ctx->kp_xy, k * 2 * sizeof(float));
        add(ctx->c_kp_count, ctx->kp_count, sizeof(int));
        if (ctx->lg_state) {
            add(ctx->c_desc, ctx->desc_f32, k * 256 * sizeof(float));
        } else {
            if (ctx->cfg.precision == 1) {
                add(ctx->c_mproj_f32, ctx->mproj_f32, k * 256 * sizeof(float));
                if (ctx->mproj_x3) add(ctx->c_mproj_x3, ctx->mproj_x3, k * 512 * sizeof(bf16));
            } else
                add(ctx->c_mproj, ctx->mproj, k * 256 * sizeof(bf16));
            add(ctx->c_mlogit, ctx->mlogit, k * sizeof(float));
        }
        g.n_arr = a;
        GNB_KERNEL(ctx, "cache_gather_kernel", cache_gather_kernel<<<dim3(16, g.n), 256, 0, ctx->stream>>>(g));
    }
    return GNB_OK;
}

// which of these rasters would be served from the cache by the next gnb_pose_candidates call with the same ids and size
// (host-side lookup, nothing is modified): lets the caller skip gathering the pixels of cached rasters
extern "C" int gnb_cache_lookup(gnb_ctx* ctx, const int64_t* tile_ids, int n_tiles, int ht, int wt, int* hit_out) {
    if (!ctx || !tile_ids || !hit_out || n_tiles < 0) return GNB_E_INVALID;
    std::vector<char> used(ctx->cache_cap, 0);
    for (int i = 0; i < n_tiles; ++i) {
        hit_out[i] = 0;
        if (ctx->cache_h != ht || ctx->cache_w != wt || tile_ids[i] < 0) continue;
        for (int e = 0; e < ctx->cache_cap; ++e)
            if (ctx->cache_ids[e] == tile_ids[i] && !used[e]) { hit_out[i] = 1; used[e] = 1; break; }
    }
    return GNB_OK;
}

static int pose_candidates_impl(gnb_ctx* ctx, const uint8_t* frame, int hq, int wq, int n_tiles, const uint8_t* const* tile_ptrs, int ht,
                                int wt, const int64_t* tile_ids, const uint8_t* dems, const double* k9, const double* affine12,
                                gnb_pose_result* results, int* n_cache_hits);

extern "C" int gnb_pose_candidates(gnb_ctx* ctx, const uint8_t* frame, int hq, int wq, int n_tiles, const uint8_t* tiles, int ht,
                                   int wt, const int64_t* tile_ids, const uint8_t* dems, const double* k9, const double* affine12,
                                   gnb_pose_result* results, int* n_cache_hits) {
    if (!ctx || !frame || !tiles || !k9 || !affine12 || !results || n_tiles < 1) return GNB_E_INVALID;
    std::vector<const uint8_t*> ptrs(n_tiles);
    for (int i = 0; i < n_tiles; ++i) ptrs[i] = tiles + (size_t)i * ht * wt;
    return pose_candidates_impl(ctx, frame, hq, wq, n_tiles, ptrs.data(), ht, wt, tile_ids, dems, k9, affine12, results, n_cache_hits);
}

// same, one pointer per raster; the pointer of a raster the cache will serve (gnb_cache_lookup) may be NULL
extern "C" int gnb_pose_candidates_ptrs(gnb_ctx* ctx, const uint8_t* frame, int hq, int wq, int n_tiles, const uint8_t* const* tile_ptrs,
                                        int ht, int wt, const int64_t* tile_ids, const uint8_t* dems, const double* k9,
                                        const double* affine12, gnb_pose_result* results, int* n_cache_hits) {
    if (!ctx || !frame || !tile_ptrs || !k9 || !affine12 || !results || n_tiles < 1) return GNB_E_INVALID;
    return pose_candidates_impl(ctx, frame, hq, wq, n_tiles, tile_ptrs, ht, wt, tile_ids, dems, k9, affine12, results, n_cache_hits);
}

static int pose_candidates_impl(gnb_ctx* ctx, const uint8_t* frame, int hq, int wq, int n_tiles, const uint8_t* const* tile_ptrs, int ht,
                                   int wt, const int64_t* tile_ids, const uint8_t* dems, const double* k9, const double* affine12,
                                   gnb_pose_result* results, int* n_cache_hits) {
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_tiles > ctx->cfg.max_batch) { GNB_SET_ERR(ctx, "n_tiles %d exceeds max_batch %d", n_tiles, ctx->cfg.max_batch); return GNB_E_CAPACITY; }
    const bool layers = ctx->lg_state != nullptr;
    if (layers && !ctx->c_desc) {
        const size_t cc = ctx->cache_cap, kk = ctx->cfg.max_keypoints;
        GNB_CUDA(ctx, cudaMalloc((void**)&ctx->c_desc, cc * kk * 256 * sizeof(float)));
        gnb_cache_clear(ctx);   // entries cached by the head-only path hold projected, not raw, descriptors
    }
    int rc;
    if ((rc = check_image(ctx, hq, wq)) || (rc = check_image(ctx, ht, wt))) return rc;
    const int sb = ctx->cfg.max_batch;
    ConvWorkspace& cw = ctx->cw;
    GnbRange range("gnb_pose_candidates");
    // cached features belong to one raster geometry: a different (ht, wt) starts from an empty cache
    if (ctx->cache_h != ht || ctx->cache_w != wt) { gnb_cache_clear(ctx); ctx->cache_h = ht; ctx->cache_w = wt; }
    // cache lookup: entry[i] = cache entry that will hold raster i's features; hit[i] = already there
    std::vector<int> entry(n_tiles), hit(n_tiles, 0);
    std::vector<char> used(ctx->cache_cap, 0);
    int hits = 0;
    for (int i = 0; i < n_tiles; ++i) {
        entry[i] = -1;
        if (tile_ids && tile_ids[i] >= 0)
            for (int e = 0; e < ctx->cache_cap; ++e)
                if (ctx->cache_ids[e] == tile_ids[i] && !used[e]) { entry[i] = e; hit[i] = 1; used[e] = 1; ++hits; break; }
    }
    for (int i = 0; i < n_tiles; ++i)
        if (!hit[i] && !tile_ptrs[i]) {
            GNB_SET_ERR(ctx, "raster %d is not in the feature cache: its pixels are required (gnb_cache_lookup)", i);
            return GNB_E_INVALID;
        }
    for (int i = 0; i < n_tiles; ++i) {
        if (entry[i] >= 0) continue;
        int best = -1;
        for (int e = 0; e < ctx->cache_cap; ++e)   // least recently used entry not claimed by this call
            if (!used[e] && (best < 0 || ctx->cache_lru[e] < ctx->cache_lru[best])) best = e;
        entry[i] = best; used[best] = 1;
        ctx->cache_ids[best] = -1;   // invalid until the raster has been extracted AND validated (committed after the final sync)
    }
    for (int i = 0; i < n_tiles; ++i) ctx->cache_lru[entry[i]] = ++ctx->cache_clock;
    if (n_cache_hits) *n_cache_hits = hits;
    // parameters: one K for all pairs
    std::vector<double> kk((size_t)n_tiles * 9);
    for (int i = 0; i < n_tiles; ++i) memcpy(&kk[(size_t)i * 9], k9, 9 * sizeof(double));
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->kmat, kk.data(), kk.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->affine, affine12, sizeof(double) * 12 * n_tiles, cudaMemcpyHostToDevice, ctx->stream));
    if (dems) GNB_CUDA(ctx, cudaMemcpyAsync(ctx->dem, dems, (size_t)n_tiles * ht * wt, cudaMemcpyHostToDevice, ctx->stream));
    // query frame -> slot 0
    GNB_CUDA(ctx, cudaMemcpyAsync(cw.img_a, frame, (size_t)hq * wq, cudaMemcpyHostToDevice, ctx->stream));
    cw.img = cw.img_a;
    if ((rc = gnb_conv_forward(ctx, 1, hq, wq, 0))) return rc;
    if ((rc = gnb_kp_select(ctx, cw.score, 1, hq, wq, 0))) return rc;
    if ((rc = gnb_describe(ctx, 1, hq, wq, 0))) return rc;
    if (!layers && (rc = gnb_match_project(ctx, 0, 1))) return rc;
    // rasters that missed the cache: one compact batch -> slots [sb, sb + m)
    std::vector<int> miss;
    for (int i = 0; i < n_tiles; ++i) if (!hit[i]) miss.push_back(i);
    const int m = (int)miss.size();
    if (m > 0) {
        for (int j = 0; j < m; ++j)
            GNB_CUDA(ctx, cudaMemcpyAsync(cw.img_b + (size_t)j * ht * wt, tile_ptrs[miss[j]], (size_t)ht * wt,
                                          cudaMemcpyHostToDevice, ctx->stream));
        cw.img = cw.img_b;
        rc = gnb_conv_forward(ctx, m, ht, wt, 0);
        cw.img = cw.img_a;
        if (rc) return rc;
        if ((rc = gnb_kp_select(ctx, cw.score, m, ht, wt, sb))) return rc;
        if ((rc = gnb_describe(ctx, m, ht, wt, sb))) return rc;
        if (!layers && (rc = gnb_match_project(ctx, sb, m))) return rc;
        for (int j = 0; j < m; ++j) cache_copy(ctx, entry[miss[j]], sb + j, true);
    }
    if ((rc = cache_gather(ctx, entry.data(), n_tiles, sb))) return rc;
    GNB_CUDA(ctx, cudaGetLastError());
    int stride_a = 0;
    if (layers) {
        // the layers refine the query features against each candidate separately: one copy of the frame per pair
        // (slots 1..n-1), then layers + head over n ordinary pairs
        const size_t kk = ctx->cfg.max_keypoints;
        for (int i = 1; i < n_tiles; ++i) {
            GNB_CUDA(ctx, cudaMemcpyAsync(ctx->desc_f32 + (size_t)i * kk * 256, ctx->desc_f32, kk * 256 * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
            GNB_CUDA(ctx, cudaMemcpyAsync(ctx->kp_xy + (size_t)i * kk * 2, ctx->kp_xy, kk * 2 * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
            GNB_CUDA(ctx, cudaMemcpyAsync(ctx->kp_count + i, ctx->kp_count, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
        }
        if ((rc = gnb_lightglue_forward(ctx, n_tiles, 0, sb, (float)hq, (float)wq, (float)ht, (float)wt))) return rc;
        if ((rc = gnb_match_project(ctx, 0, n_tiles))) return rc;
        if ((rc = gnb_match_project(ctx, sb, n_tiles))) return rc;
        stride_a = 1;
    }
    if ((rc = gnb_match_pairs(ctx, n_tiles, 0, sb, stride_a))) return rc;
    if ((rc = gnb_pnp_pairs(ctx, n_tiles, ht, wt, dems ? 1 : 0, ht, wt, 1, ctx->cfg.min_matches, 1, stride_a))) return rc;
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->out_host, ctx->out_dev, sizeof(PairOut) * n_tiles, cudaMemcpyDeviceToHost, ctx->stream));
    GNB_SYNC(ctx);
    int err = GNB_OK;
    for (int b = 0; b < n_tiles; ++b) {
        pairout_to_result(ctx->out_host[b], &results[b]);
        if (results[b].n_kp_qry < 0 || results[b].n_kp_ref < 0) { GNB_SET_ERR(ctx, "NMS candidate buffer overflow"); err = GNB_E_CAPACITY; continue; }
        // the raster's features are valid: only now does its cache entry get its id (an entry claimed above stays
        // invalid on every error path, so a later call can never hit half-written features)
        if (!hit[b] && tile_ids && tile_ids[b] >= 0) ctx->cache_ids[entry[b]] = tile_ids[b];
        if (results[b].status == GNB_E_RANGE && !err) { GNB_SET_ERR(ctx, "reference keypoint outside the DEM in candidate %d", b); err = GNB_E_RANGE; }
    }
    return err;
}

// ------------------------------------------------------------------------------------------------
// stage-isolated hooks
extern "C" int gnb_dense(gnb_ctx* ctx, const uint8_t* image, int h, int w, int stride, float* out_score, float* out_dense) {
    if (!ctx || !image) return GNB_E_INVALID;
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = check_image(ctx, h, w))) return rc;
    GNB_CUDA(ctx, cudaMemcpy2DAsync(ctx->cw.img, w, image, stride, w, h, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = gnb_conv_forward(ctx, 1, h, w, 1))) return rc;
    if (out_score) GNB_CUDA(ctx, cudaMemcpyAsync(out_score, ctx->cw.score, sizeof(float) * h * w, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_dense) GNB_CUDA(ctx, cudaMemcpyAsync(out_dense, ctx->cw.dense, sizeof(float) * (h / 8) * (w / 8) * 256, cudaMemcpyDeviceToHost, ctx->stream));
    GNB_SYNC(ctx);
    return GNB_OK;
}

int gnb_split_to_f32(gnb_ctx* ctx, const bf16* in, float* out, size_t pixels, int c);   // conv_x3.cu

__global__ void bf16_to_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __bfloat162float(in[i]);
}

// run the dense stack (and optionally K2 + K3 into slots [0, n)) on a batch of n images: parity hook for the
// persistent multi-tile / multi-image loops of the conv kernels
extern "C" int gnb_dense_batch(gnb_ctx* ctx, const uint8_t* images, int n, int h, int w, int dense_desc, int with_keypoints) {
    if (!ctx || !images || n < 1) return GNB_E_INVALID;
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = check_image(ctx, h, w))) return rc;
    if (n > ctx->cfg.max_batch) { GNB_SET_ERR(ctx, "batch %d exceeds max_batch %d", n, ctx->cfg.max_batch); return GNB_E_CAPACITY; }
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->cw.img, images, (size_t)n * h * w, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = gnb_conv_forward(ctx, n, h, w, dense_desc))) return rc;
    if (with_keypoints) {
        if ((rc = gnb_kp_select(ctx, ctx->cw.score, n, h, w, 0))) return rc;
        if ((rc = gnb_describe(ctx, n, h, w, 0))) return rc;
    }
    GNB_SYNC(ctx);
    return GNB_OK;
}

// keypoints + descriptors of keypoint slot `slot` (after gnb_dense_batch(with_keypoints) / gnb_pose_batch)
extern "C" int gnb_slot_keypoints(gnb_ctx* ctx, int slot, float* out_xy, float* out_score, float* out_desc, int cap, int* n_out) {
    if (!ctx || !n_out || slot < 0 || slot >= ctx->kp_slots) return GNB_E_INVALID;
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    int n = 0, rc;
    if ((rc = read_count(ctx, ctx->kp_count + slot, &n))) return rc;
    if (n < 0) { GNB_SET_ERR(ctx, "NMS candidate buffer overflow"); return GNB_E_CAPACITY; }
    n = n < cap ? n : cap;
    *n_out = n;
    const size_t k = ctx->cfg.max_keypoints;
    if (n > 0) {
        if (out_xy) GNB_CUDA(ctx, cudaMemcpyAsync(out_xy, ctx->kp_xy + slot * k * 2, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, ctx->stream));
        if (out_score) GNB_CUDA(ctx, cudaMemcpyAsync(out_score, ctx->kp_score + slot * k, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
        if (out_desc) GNB_CUDA(ctx, cudaMemcpyAsync(out_desc, ctx->desc_f32 + slot * k * 256, sizeof(float) * 256 * n, cudaMemcpyDeviceToHost, ctx->stream));
        GNB_SYNC(ctx);
    }
    return GNB_OK;
}

// match index pairs (query keypoint, reference keypoint) of pair `pair` of the last batch / matcher call
extern "C" int gnb_pair_matches(gnb_ctx* ctx, int pair, int32_t* out_idx, int cap, int* n_out) {
    if (!ctx || !n_out || pair < 0 || pair >= ctx->cfg.max_batch) return GNB_E_INVALID;
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    int n = 0, rc;
    if ((rc = read_count(ctx, ctx->match_count + pair, &n))) return rc;
    n = n < cap ? n : cap;
    *n_out = n;
    if (n > 0 && out_idx) {
        GNB_CUDA(ctx, cudaMemcpyAsync(out_idx, ctx->match_idx + (size_t)pair * ctx->cfg.max_keypoints * 2, sizeof(int) * 2 * n, cudaMemcpyDeviceToHost, ctx->stream));
        GNB_SYNC(ctx);
    }
    return GNB_OK;
}

extern "C" int gnb_layer_activation_at(gnb_ctx* ctx, const char* layer, int image_index, float* out, size_t out_floats);
extern "C" int gnb_layer_activation(gnb_ctx* ctx, const char* layer, float* out, size_t out_floats) {
    return gnb_layer_activation_at(ctx, layer, 0, out, out_floats);
}

extern "C" int gnb_layer_activation_at(gnb_ctx* ctx, const char* layer, int image_index, float* out, size_t out_floats) {
    if (!ctx || !layer || !out) return GNB_E_INVALID;
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    const ConvWorkspace& cw = ctx->cw;
    if (image_index < 0 || image_index >= cw.n) { GNB_SET_ERR(ctx, "image index %d outside the last pass (%d images)", image_index, cw.n); return GNB_E_INVALID; }
    const size_t ii = (size_t)image_index;
    struct { const char* name; const bf16* p; int div, c; } tbl[] = {
        {"conv1a", cw.a1a, 1, 64}, {"pool1", cw.p1, 2, 64},   {"conv2a", cw.a2a, 2, 64}, {"pool2", cw.p2, 4, 64},
        {"conv3a", cw.a3a, 4, 128}, {"pool3", cw.p3, 8, 128}, {"conv4a", cw.a4a, 8, 128}, {"conv4b", cw.a4b, 8, 128},
        {"convPa", cw.apa, 8, 256}, {"convDa", cw.ada, 8, 256},
    };
    for (auto& t : tbl) {
        if (strcmp(t.name, layer) == 0) {
            const size_t n = (size_t)(cw.h / t.div) * (cw.w / t.div) * t.c;
            if (n != out_floats) { GNB_SET_ERR(ctx, "layer %s has %zu floats, caller passed %zu", layer, n, out_floats); return GNB_E_INVALID; }
            int rc;
            if ((rc = gnb_ensure_stage(ctx, n, 0))) return rc;
            if (ctx->cfg.precision == 1) {
                if ((rc = gnb_split_to_f32(ctx, t.p + ii * 2 * n, ctx->stage_a, n / t.c, t.c))) return rc;
                GNB_CUDA(ctx, cudaMemcpyAsync(out, ctx->stage_a, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
                GNB_SYNC(ctx);
                return GNB_OK;
            }
            GNB_KERNEL(ctx, "bf16_to_f32_kernel", bf16_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(t.p + ii * n, ctx->stage_a, n));
            GNB_CUDA(ctx, cudaMemcpyAsync(out, ctx->stage_a, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
            GNB_SYNC(ctx);
            return GNB_OK;
        }
    }
    if (strcmp(layer, "semi") == 0) {
        const size_t n = (size_t)(cw.h / 8) * (cw.w / 8) * 65;
        if (n != out_floats) return GNB_E_INVALID;
        GNB_CUDA(ctx, cudaMemcpyAsync(out, cw.semi + ii * n, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        GNB_SYNC(ctx);
        return GNB_OK;
    }
    if (strcmp(layer, "score") == 0 || strcmp(layer, "dense") == 0) {
        const bool sc = layer[0] == 's';
        const size_t n = sc ? (size_t)cw.h * cw.w : (size_t)(cw.h / 8) * (cw.w / 8) * 256;
        if (n != out_floats) { GNB_SET_ERR(ctx, "layer %s has %zu floats, caller passed %zu", layer, n, out_floats); return GNB_E_INVALID; }
        GNB_CUDA(ctx, cudaMemcpyAsync(out, (sc ? cw.score : cw.dense) + ii * n, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        GNB_SYNC(ctx);
        return GNB_OK;
    }
    GNB_SET_ERR(ctx, "unknown layer '%s'", layer);
    return GNB_E_INVALID;
}

extern "C" int gnb_select_keypoints(gnb_ctx* ctx, const float* score, int h, int w, float* out_xy, float* out_score,
                                    int cap, int* n_out) {
    if (!ctx || !score || !n_out) return GNB_E_INVALID;
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (h <= 0 || w <= 0 || (size_t)h * w > ctx->cw.cap_pixels) { GNB_SET_ERR(ctx, "score map exceeds the workspace"); return GNB_E_CAPACITY; }
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->cw.score, score, sizeof(float) * h * w, cudaMemcpyHostToDevice, ctx->stream));
    int rc;
    if ((rc = gnb_kp_select(ctx, ctx->cw.score, 1, h, w, 0))) return rc;
    int n = 0;
    if ((rc = read_count(ctx, ctx->kp_count, &n))) return rc;
    if (n < 0) { GNB_SET_ERR(ctx, "NMS candidate buffer overflow"); return GNB_E_CAPACITY; }
    n = n < cap ? n : cap;
    *n_out = n;
    if (n > 0) {
        if (out_xy) GNB_CUDA(ctx, cudaMemcpyAsync(out_xy, ctx->kp_xy, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, ctx->stream));
        if (out_score) GNB_CUDA(ctx, cudaMemcpyAsync(out_score, ctx->kp_score, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
        GNB_SYNC(ctx);
    }
    return GNB_OK;
}

extern "C" int gnb_sample_descriptors(gnb_ctx* ctx, const float* dense, int hc, int wc, const float* xy, int n, int img_h,
                                      int img_w, float* out_desc) {
    if (!ctx || !dense || (n > 0 && (!xy || !out_desc))) return GNB_E_INVALID;
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n > ctx->cfg.max_keypoints || (size_t)hc * wc * 64 > ctx->cw.cap_pixels) { GNB_SET_ERR(ctx, "exceeds workspace"); return GNB_E_CAPACITY; }
    if (n == 0) return GNB_OK;
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->cw.dense, dense, sizeof(float) * hc * wc * 256, cudaMemcpyHostToDevice, ctx->stream));
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->kp_xy, xy, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->kp_count, &n, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    GNB_SYNC(ctx);
    // the sampler derives hc,wc from the image size; honour the caller's explicit map size
    if (hc != img_h / 8 || wc != img_w / 8) { GNB_SET_ERR(ctx, "dense map must be image/8"); return GNB_E_INVALID; }
    int rc;
    if ((rc = gnb_kp_sample(ctx, ctx->cw.dense, 1, img_h, img_w, 0))) return rc;
    GNB_CUDA(ctx, cudaMemcpyAsync(out_desc, ctx->desc_f32, sizeof(float) * 256 * n, cudaMemcpyDeviceToHost, ctx->stream));
    GNB_SYNC(ctx);
    return GNB_OK;
}

// full score matrix for small problems: one thread per (i, j), direct dot product of the
// projected descriptors + the LSE vectors of pass 0.
__global__ void score_matrix_kernel(const bf16* __restrict__ ma, const bf16* __restrict__ mb, const float* __restrict__ lse_a,
                                    const float* __restrict__ lse_b, const float* __restrict__ la,
                                    const float* __restrict__ lb, int na, int nb, float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= nb || i >= na) return;
    float s = 0.f;
    for (int k = 0; k < 256; ++k) s = fmaf(__bfloat162float(ma[(size_t)i * 256 + k]), __bfloat162float(mb[(size_t)j * 256 + k]), s);
    out[(size_t)i * nb + j] = __fadd_rn(__fadd_rn(__fadd_rn(__fsub_rn(s, lse_a[i]), __fsub_rn(s, lse_b[j])), la[i]), lb[j]);
}

extern "C" int gnb_match_scores(gnb_ctx* ctx, const float* desc_a, int n_a, const float* desc_b, int n_b, float* out_scores) {
    if (!ctx || !desc_a || !desc_b || !out_scores || n_a < 1 || n_b < 1) return GNB_E_INVALID;
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->cfg.precision == 1) { GNB_SET_ERR(ctx, "gnb_match_scores is a hook of the bf16 matcher (precision = 0)"); return GNB_E_INVALID; }
    int rc;
    if ((rc = load_descs(ctx, desc_a, n_a, desc_b, n_b, 0))) return rc;
    if ((rc = gnb_match_pairs(ctx, 1, 0, ctx->cfg.max_batch))) return rc;  // fills row_lse for both sides
    if ((rc = gnb_ensure_stage(ctx, (size_t)n_a * n_b, 0))) return rc;
    const int k = ctx->cfg.max_keypoints, sb = ctx->cfg.max_batch;
    dim3 grid(ceil_div(n_b, 128), n_a);
    GNB_KERNEL(ctx, "score_matrix_kernel", score_matrix_kernel<<<grid, 128, 0, ctx->stream>>>(ctx->mproj, ctx->mproj + (size_t)sb * k * 256, ctx->row_lse,
                                                       ctx->row_lse + (size_t)sb * k, ctx->mlogit, ctx->mlogit + (size_t)sb * k,
                                                       n_a, n_b, ctx->stage_a));
    GNB_CUDA(ctx, cudaMemcpyAsync(out_scores, ctx->stage_a, sizeof(float) * n_a * n_b, cudaMemcpyDeviceToHost, ctx->stream));
    GNB_SYNC(ctx);
    return GNB_OK;
}

extern "C" int gnb_ransac_debug(gnb_ctx* ctx, int32_t* out_counts, float* out_hyp, int* out_best) {
    if (!ctx) return GNB_E_INVALID;
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int it = ctx->cfg.ransac_iters;
    if (out_counts) GNB_CUDA(ctx, cudaMemcpyAsync(out_counts, ctx->hyp_count, sizeof(int) * it, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_hyp) GNB_CUDA(ctx, cudaMemcpyAsync(out_hyp, ctx->hyp, sizeof(float) * 12 * it, cudaMemcpyDeviceToHost, ctx->stream));
    GNB_SYNC(ctx);
    if (out_best) *out_best = ctx->out_host[0].best_hypothesis;
    return GNB_OK;
}
