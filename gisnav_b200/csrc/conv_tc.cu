// conv_tc.cu — K1 on the 5th-generation tensor cores: 3x3 / 1x1 convolutions of the dense stack as
// implicit GEMMs.  M = 128 output pixels (an 8 x 16 spatial tile = 128 TMEM lanes), N = Cout,
// K = taps x Cin.  Activations are bf16 NHWC, so one pixel's 64 channels are one 128-byte row of the
// K-major 128B-swizzled operand layout; a 4-D TMA box (64 ch, 16 px, 8 rows, 1 image) at a
// tap-shifted coordinate IS the A operand of that tap, and TMA's out-of-bounds zero fill is the
// convolution's zero padding.  Weights are bf16 [tap][Cout][Cin] (K-major B operand).
// fp32 accumulators live in TMEM; the epilogue (bias, ReLU, 2x2 max-pool by warp shuffles, bf16
// pack) reads them with tcgen05.ld, one pixel per thread.
//
//   warp 0  TMA producer     warp 1  MMA issuer     warp 2  TMEM allocator     warps 4-7  epilogue
#include "tc_common.cuh"

#define CT_TH 8
#define CT_TW 16
#define CT_A_BYTES (128 * 128)   // 128 pixels x 64 ch x 2 B
#define CT_STAGES 4

int* gnb_tc_err_dev(gnb_ctx* ctx);

struct ConvTcLayerMaps { CUtensorMap w; int valid; };
static ConvTcLayerMaps g_wmaps[GNB_NUM_LAYERS];

template <int NPAD>
__global__ void __launch_bounds__(256) conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_in,
                                                      const __grid_constant__ CUtensorMap tmap_w,
                                                      const float* __restrict__ bias, int h, int w, int cin, int ks, int cout,
                                                      bf16* __restrict__ out_bf, float* __restrict__ out_f, int relu, int pool,
                                                      int* err) {
    constexpr int B_BYTES = NPAD * 128;
    constexpr int STAGE_BYTES = CT_A_BYTES + B_BYTES;
    constexpr int TMEM_COLS = NPAD <= 32 ? 32 : NPAD <= 64 ? 64 : NPAD <= 128 ? 128 : 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CT_STAGES * STAGE_BYTES);
    uint64_t* full = bars;                 // [CT_STAGES]
    uint64_t* empty = bars + CT_STAGES;    // [CT_STAGES]
    uint64_t* acc_full = bars + 2 * CT_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * CT_STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img = blockIdx.z, y0 = blockIdx.y * CT_TH, x0 = blockIdx.x * CT_TW;
    const int r = ks / 2, taps = ks * ks, kchunks = cin / 64;
    const int n_iters = taps * kchunks;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmap_in);
        tc::tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < CT_STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(acc_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) {
        tc::tmem_alloc(tmem_slot, TMEM_COLS);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < n_iters; ++it) {
                const int s = it % CT_STAGES;
                const uint32_t ph = (it / CT_STAGES) & 1;
                if (it >= CT_STAGES && !tc::mbar_wait(&empty[s], ph ^ 1, err, 201)) break;
                const int tap = it / kchunks, kc = it % kchunks;
                const int dy = tap / ks - r, dx = tap % ks - r;
                uint8_t* sa = smem + s * STAGE_BYTES;
                tc::mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
                tc::tma_load_4d(sa, &tmap_in, &full[s], kc * 64, x0 + dx, y0 + dy, img);
                tc::tma_load_3d(sa + CT_A_BYTES, &tmap_w, &full[s], kc * 64, 0, tap);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(128, NPAD);
            int it = 0;
            for (; it < n_iters; ++it) {
                const int s = it % CT_STAGES;
                const uint32_t ph = (it / CT_STAGES) & 1;
                if (!tc::mbar_wait(&full[s], ph, err, 202)) break;
                tc::tc_fence_after();
                const uint32_t a_addr = tc::smem_u32(smem + s * STAGE_BYTES);
                const uint32_t b_addr = a_addr + CT_A_BYTES;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t da = tc::make_smem_desc_sw128(a_addr + k * 32, 1024);
                    const uint64_t db = tc::make_smem_desc_sw128(b_addr + k * 32, 1024);
                    tc::umma_bf16(tmem_base, da, db, idesc, (it | k) ? 1u : 0u);
                }
                tc::umma_commit(&empty[s]);
            }
            tc::umma_commit(acc_full);
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int yl = 2 * q + (lane >> 4), xl = lane & 15;
        const int y = y0 + yl, x = x0 + xl;
        const bool ok = tc::mbar_wait(acc_full, 0, err, 203);
        tc::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool inside = (y < h) && (x < w);
        size_t pix;
        bool writer;
        if (pool) {
            writer = inside && ((xl | yl) & 1) == 0;
            pix = ((size_t)img * (h / 2) + (y >> 1)) * (size_t)(w / 2) + (x >> 1);
        } else {
            writer = inside;
            pix = ((size_t)img * h + y) * (size_t)w + x;
        }
        if (ok) {
#pragma unroll 1
            for (int c0 = 0; c0 < NPAD; c0 += 32) {
                uint32_t v[32];
                tc::tmem_ld32(taddr + c0, v);
                tc::tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float a = __uint_as_float(v[i]) + __ldg(&bias[c0 + i]);
                    if (relu) a = fmaxf(a, 0.f);
                    if (pool) {
                        a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, 1));
                        a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, 16));
                    }
                    f[i] = a;
                }
                if (writer) {
                    if (out_bf) {
                        bf16* o = out_bf + pix * cout + c0;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (c0 + g * 8 < cout) {
                                __align__(16) __nv_bfloat162 p[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[g * 8 + 2 * i], f[g * 8 + 2 * i + 1]);
                                *reinterpret_cast<uint4*>(o + g * 8) = *reinterpret_cast<const uint4*>(p);
                            }
                        }
                    } else {
                        float* o = out_f + pix * cout + c0;
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (c0 + i < cout) o[i] = f[i];
                    }
                }
            }
        }
        tc::tc_fence_before();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int NPAD>
static int launch_conv_tc(gnb_ctx* ctx, const CUtensorMap& tin, const CUtensorMap& tw, const ConvLayer& L, int n, int h, int w,
                          bf16* out_bf, float* out_f, int relu, int pool) {
    constexpr int smem = CT_STAGES * (CT_A_BYTES + NPAD * 128) + 1024 + 256;
    static bool attr_set = false;
    if (!attr_set) {
        GNB_CUDA(ctx, cudaFuncSetAttribute(conv_tc_kernel<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    dim3 grid(ceil_div(w, CT_TW), ceil_div(h, CT_TH), n);
    GNB_KERNEL(ctx, "conv_tc_kernel", conv_tc_kernel<NPAD><<<grid, 256, smem, ctx->stream>>>(
        tin, tw, L.bias, h, w, L.cin, L.ks, L.cout, out_bf, out_f, relu, pool, gnb_tc_err_dev(ctx)));
    return GNB_OK;
}

int gnb_conv_tc_init(gnb_ctx* ctx) {
    if (!gnb_tc_err_dev(ctx)) { GNB_SET_ERR(ctx, "cannot allocate the host-mapped error word"); return GNB_E_CUDA; }
    for (int l = 0; l < GNB_NUM_LAYERS; ++l) {
        const ConvLayer& L = ctx->layers[l];
        g_wmaps[l].valid = 0;
        if (L.cin % 64) continue;  // conv1a (Cin = 1) stays on the CUDA cores
        const uint64_t dims[3] = {(uint64_t)L.cin, (uint64_t)L.cout_pad, (uint64_t)(L.ks * L.ks)};
        const uint64_t strides[2] = {(uint64_t)L.cin * 2, (uint64_t)L.cout_pad * L.cin * 2};
        const uint32_t box[3] = {64, (uint32_t)L.cout_pad, 1};
        int rc = gnb_make_tmap_bf16(ctx, &g_wmaps[l].w, L.w, 3, dims, strides, box);
        if (rc) return rc;
        g_wmaps[l].valid = 1;
    }
    return GNB_OK;
}

int gnb_conv_tc_layer(gnb_ctx* ctx, const ConvLayer& L, const bf16* in, int n, int h, int w, bf16* out_bf, float* out_f,
                      int relu, int pool) {
    const int lid = (int)(&L - ctx->layers);
    if (lid < 0 || lid >= GNB_NUM_LAYERS || !g_wmaps[lid].valid) return GNB_E_INVALID;
    if (pool && ((h | w) & 1)) return GNB_E_INVALID;
    CUtensorMap tin;
    const uint64_t dims[4] = {(uint64_t)L.cin, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)L.cin * 2, (uint64_t)w * L.cin * 2, (uint64_t)h * w * L.cin * 2};
    const uint32_t box[4] = {64, CT_TW, CT_TH, 1};
    int rc = gnb_make_tmap_bf16(ctx, &tin, const_cast<bf16*>(in), 4, dims, strides, box);
    if (rc) return rc;
    switch (L.cout_pad) {
        case 64: return launch_conv_tc<64>(ctx, tin, g_wmaps[lid].w, L, n, h, w, out_bf, out_f, relu, pool);
        case 96: return launch_conv_tc<96>(ctx, tin, g_wmaps[lid].w, L, n, h, w, out_bf, out_f, relu, pool);
        case 128: return launch_conv_tc<128>(ctx, tin, g_wmaps[lid].w, L, n, h, w, out_bf, out_f, relu, pool);
        case 256: return launch_conv_tc<256>(ctx, tin, g_wmaps[lid].w, L, n, h, w, out_bf, out_f, relu, pool);
        default: return GNB_E_INVALID;
    }
}
