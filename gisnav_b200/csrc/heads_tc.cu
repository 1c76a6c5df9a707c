// heads_tc.cu — the two 1x1 heads of the dense stack on tcgen05, each fused with what follows it.
//
// score head  : convPb (256 -> 65, 1x1) + softmax over the 65 logits + dustbin drop + 8x8
//               depth-to-space, written straight into the full-resolution score map.  Persistent
//               CTAs, weights (48 KB) resident in shared memory, A = 8 x 16 cells x 256 ch per tile
//               (four 128B-swizzled TMA boxes), two TMEM accumulator stages.
// descriptor head : convDb (256 -> 256, 1x1) evaluated ONLY at the four coarse cells around each
//               selected keypoint (on demand) + per-cell L2 normalisation + bilinear interpolation +
//               final L2 normalisation (K3, oracle/sample_ref.py).  The dense 256-channel fp32 map
//               (1 KB per cell) is never written.  M = 128 rows = 32 keypoints x 4 corners; the A
//               operand is gathered by the CTA's threads into the 128B-swizzle layout
//               (16-byte chunk index XOR row%8), then tcgen05.mma with the resident 128 KB weights.
#include "tc_common.cuh"

#include <math.h>

int* gnb_tc_err_dev(gnb_ctx* ctx);

// ------------------------------------------------------------------------------------------------
// score head
#define SH_NPAD 96
#define SH_A_BYTES (4 * 128 * 128)        // 4 chunks x 128 cells x 128 B = 64 KB
#define SH_W_BYTES (4 * SH_NPAD * 128)    // 48 KB
#define SH_STAGES 2
#define SH_SMEM (1024 + SH_W_BYTES + SH_STAGES * SH_A_BYTES + 256 + SH_NPAD * 4)
#define SH_SMEM_X3 (1024 + 2 * SH_W_BYTES + SH_STAGES * SH_A_BYTES + 256 + SH_NPAD * 4)

// X3 (fp32-faithful mode): the activation is [hi: 256 | lo: 256] per cell and the weights [hi | lo] per output row (96 KB
// resident).  A tile is fed as TWO ring stages — the four hi chunks, then the four lo chunks — and every hi chunk is
// multiplied with the hi and the lo weight block, every lo chunk with the hi block: three MMAs per product into one
// accumulator, then the same softmax / depth-to-space epilogue on the fp32 logits.
template <bool X3>
__global__ void __launch_bounds__(256, 1) score_head_kernel(const __grid_constant__ CUtensorMap tmap_in,
                                                            const __grid_constant__ CUtensorMap tmap_w,
                                                            const float* __restrict__ bias, int hc, int wc, int n_img,
                                                            float* __restrict__ score, int* err) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int WB = (X3 ? 2 : 1) * SH_W_BYTES, HALVES = X3 ? 2 : 1;
    uint8_t* sW = smem;
    uint8_t* sA = smem + WB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + SH_STAGES * SH_A_BYTES);
    uint64_t* w_full = bars;
    uint64_t* a_full = bars + 1;
    uint64_t* a_empty = a_full + SH_STAGES;
    uint64_t* t_full = a_empty + SH_STAGES;
    uint64_t* t_empty = t_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
    float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (wc + 15) / 16, tiles_y = (hc + 7) / 8;
    const int tiles_per_img = tiles_x * tiles_y, total = tiles_per_img * n_img;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmap_in);
        tc::tma_prefetch_desc(&tmap_w);
        tc::mbar_init(w_full, 1);
        for (int s = 0; s < SH_STAGES; ++s) { tc::mbar_init(&a_full[s], 1); tc::mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&t_full[s], 1); tc::mbar_init(&t_empty[s], 4); }
        tc::fence_barrier_init();
    }
    if (warp == 2) { tc::tmem_alloc(tmem_slot, 256); tc::tmem_relinquish(); }
    if (threadIdx.x >= 128 && threadIdx.x - 128 < SH_NPAD) s_bias[threadIdx.x - 128] = bias[threadIdx.x - 128];
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(w_full, WB);
            for (int c = 0; c < 4 * HALVES; ++c) tc::tma_load_3d(sW + c * SH_NPAD * 128, &tmap_w, w_full, c * 64, 0, 0);
            int i = 0;   // ring index over (tile, half)
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int img = tile / tiles_per_img, rem = tile % tiles_per_img;
                const int y0 = (rem / tiles_x) * 8, x0 = (rem % tiles_x) * 16;
                bool ok = true;
                for (int hf = 0; hf < HALVES; ++hf, ++i) {
                    const int s = i % SH_STAGES;
                    if (i >= SH_STAGES && !tc::mbar_wait(&a_empty[s], ((i / SH_STAGES) & 1) ^ 1, err, 301)) { ok = false; break; }
                    tc::mbar_arrive_expect_tx(&a_full[s], SH_A_BYTES);
                    for (int c = 0; c < 4; ++c)
                        tc::tma_load_4d(sA + s * SH_A_BYTES + c * 128 * 128, &tmap_in, &a_full[s], hf * 256 + c * 64, x0, y0, img);
                }
                if (!ok) break;
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = tc::make_idesc_bf16(128, SH_NPAD);
        bool ok = tc::mbar_wait(w_full, 0, err, 302);
        const uint64_t db0 = tc::make_smem_desc_sw128(tc::smem_u32(sW), 1024);
        const uint64_t da00 = tc::make_smem_desc_sw128(tc::smem_u32(sA), 1024);
        int i = 0, ti = 0;
        for (int tile = blockIdx.x; ok && tile < total; tile += gridDim.x, ++ti) {
            const int as = ti & 1;
            if (ti >= 2 && !tc::mbar_wait(&t_empty[as], ((ti >> 1) & 1) ^ 1, err, 304)) break;
            const uint32_t d_tmem = tmem_base + (uint32_t)(as * 128);
#pragma unroll
            for (int hf = 0; hf < HALVES; ++hf, ++i) {
                const int s = i % SH_STAGES;
                if (!tc::mbar_wait(&a_full[s], (i / SH_STAGES) & 1, err, 303)) { ok = false; break; }
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint64_t da0 = da00 + (uint64_t)((s * SH_A_BYTES) >> 4);
#pragma unroll
                    for (int term = 0; term < HALVES; ++term) {
                        if (term < (hf == 0 ? HALVES : 1)) {   // hi half: weight blocks hi (and lo); lo half: weight block hi
#pragma unroll
                            for (int c = 0; c < 4; ++c)
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    tc::umma_bf16(d_tmem, da0 + (uint64_t)((c * 128 * 128 + k * 32) >> 4),
                                                  db0 + (uint64_t)(((term * 4 + c) * SH_NPAD * 128 + k * 32) >> 4), idesc, (hf | term | c | k) ? 1u : 0u);
                        }
                    }
                    tc::umma_commit(&a_empty[s]);
                    if (hf == HALVES - 1) tc::umma_commit(&t_full[as]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int yl = 2 * q + (lane >> 4), xl = lane & 15;
        const int w_full_res = wc * 8;
        int i = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
            const int as = i & 1;
            const int img = tile / tiles_per_img, rem = tile % tiles_per_img;
            const int cy = (rem / tiles_x) * 8 + yl, cx = (rem % tiles_x) * 16 + xl;
            if (!tc::mbar_wait(&t_full[as], (i >> 1) & 1, err, 305)) break;
            tc::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 128);
            float v[65];
            {
                uint32_t r[32];
                tc::tmem_ld32(taddr, r);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + s_bias[j];
                tc::tmem_ld32(taddr + 32, r);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[32 + j] = __uint_as_float(r[j]) + s_bias[32 + j];
                tc::tmem_ld32(taddr + 64, r);
                tc::tmem_ld_wait();
                v[64] = __uint_as_float(r[0]) + s_bias[64];
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&t_empty[as]);   // accumulator drained: the next tile may reuse it
            // same operation order as softmax_d2s_kernel
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < 65; ++j) m = fmaxf(m, v[j]);
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 65; ++j) { v[j] = expf(v[j] - m); sum += v[j]; }
            const float inv = 1.0f / sum;
            if (cy < hc && cx < wc) {
                float* o = score + ((size_t)img * hc * 8 + (size_t)cy * 8) * w_full_res + (size_t)cx * 8;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    *reinterpret_cast<float4*>(o + (size_t)r * w_full_res) =
                        make_float4(v[r * 8 + 0] * inv, v[r * 8 + 1] * inv, v[r * 8 + 2] * inv, v[r * 8 + 3] * inv);
                    *reinterpret_cast<float4*>(o + (size_t)r * w_full_res + 4) =
                        make_float4(v[r * 8 + 4] * inv, v[r * 8 + 5] * inv, v[r * 8 + 6] * inv, v[r * 8 + 7] * inv);
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 256); }
}

int gnb_score_head_tc(gnb_ctx* ctx, const CUtensorMap* tmap_w, const bf16* apa, const float* bias, int n, int hc, int wc,
                      float* score) {
    const bool x3 = ctx->cfg.precision == 1;
    const uint64_t cin = x3 ? 512 : 256;   // fp32-faithful mode: [hi: 256 | lo: 256] per cell
    CUtensorMap tin;
    const uint64_t dims[4] = {cin, (uint64_t)wc, (uint64_t)hc, (uint64_t)n};
    const uint64_t strides[3] = {cin * 2, (uint64_t)wc * cin * 2, (uint64_t)hc * wc * cin * 2};
    const uint32_t box[4] = {64, 16, 8, 1};
    int rc = gnb_make_tmap_bf16(ctx, &tin, const_cast<bf16*>(apa), 4, dims, strides, box);
    if (rc) return rc;
    const int total = ceil_div(wc, 16) * ceil_div(hc, 8) * n;
    const int grid = total < ctx->sm_count ? total : ctx->sm_count;
    if (x3) {
        GNB_CUDA(ctx, gnb_func_smem(ctx, score_head_kernel<true>, SH_SMEM_X3));
        GNB_KERNEL(ctx, "score_head_x3", score_head_kernel<true><<<grid, 256, SH_SMEM_X3, ctx->stream>>>(tin, *tmap_w, bias, hc, wc, n, score, gnb_tc_err_dev(ctx)));
    } else {
        GNB_CUDA(ctx, gnb_func_smem(ctx, score_head_kernel<false>, SH_SMEM));
        GNB_KERNEL(ctx, "score_head_tc", score_head_kernel<false><<<grid, 256, SH_SMEM, ctx->stream>>>(tin, *tmap_w, bias, hc, wc, n, score, gnb_tc_err_dev(ctx)));
    }
    return GNB_OK;
}

// ------------------------------------------------------------------------------------------------
// descriptor head on demand
#define DH_W_BYTES (4 * 256 * 128)   // 128 KB: four K chunks of [256 out rows x 128 B]
#define DH_A_BYTES (4 * 128 * 128)   // 64 KB: four K chunks of [128 gathered rows x 128 B]
#define DH_SMEM (1024 + DH_W_BYTES + DH_A_BYTES + 256 + 256 * 4)

__global__ void __launch_bounds__(256, 1) desc_head_kernel(const __grid_constant__ CUtensorMap tmap_w, const bf16* __restrict__ ada,
                                                           const float* __restrict__ bias, int hc, int wc, int img_h, int img_w,
                                                           const float* __restrict__ kp_xy, const int* __restrict__ kp_count,
                                                           int slot0, int k_cap, float* __restrict__ desc, int* err) {
    const int b = blockIdx.y, slot = slot0 + b;
    const int n_kp = max(kp_count[slot], 0);
    const int kp0 = blockIdx.x * 32;
    if (kp0 >= n_kp) return;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW = smem;
    uint8_t* sA = smem + DH_W_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + DH_A_BYTES);
    uint64_t* w_full = bars;
    uint64_t* acc_full = bars + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmap_w);
        tc::mbar_init(w_full, 1);
        tc::mbar_init(acc_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) { tc::tmem_alloc(tmem_slot, 256); tc::tmem_relinquish(); }
    s_bias[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    if (warp == 0 && lane == 0) {
        tc::mbar_arrive_expect_tx(w_full, DH_W_BYTES);
        for (int c = 0; c < 4; ++c) tc::tma_load_3d(sW + c * 256 * 128, &tmap_w, w_full, c * 64, 0, 0);
    }
    // ---- gather: row m = 4 * (keypoint - kp0) + corner; 512 B per row = 4 chunks x 8 sixteen-byte pieces
    const bf16* base = ada + (size_t)b * hc * wc * 256;
    for (int item = threadIdx.x; item < 128 * 32; item += 256) {
        const int m = item >> 5, piece = item & 31;       // piece = chunk * 8 + j
        const int kp = kp0 + (m >> 2), corner = m & 3;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (kp < n_kp) {
            const float x = kp_xy[((size_t)slot * k_cap + kp) * 2 + 0], y = kp_xy[((size_t)slot * k_cap + kp) * 2 + 1];
            const float gx = __fdiv_rn(__fsub_rn(x, 3.5f), (float)img_w - 4.5f);
            const float gy = __fdiv_rn(__fsub_rn(y, 3.5f), (float)img_h - 4.5f);
            const int cx = (int)floorf(__fmul_rn(gx, (float)(wc - 1))) + (corner & 1);
            const int cy = (int)floorf(__fmul_rn(gy, (float)(hc - 1))) + (corner >> 1);
            if (cx >= 0 && cx < wc && cy >= 0 && cy < hc)
                val = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)cy * wc + cx) * 256) + piece);
        }
        const int chunk = piece >> 3, j = piece & 7;
        *reinterpret_cast<uint4*>(sA + chunk * 128 * 128 + m * 128 + ((j ^ (m & 7)) << 4)) = val;
    }
    tc::fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async proxy
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 1) {
        const bool ok = tc::mbar_wait(w_full, 0, err, 311);
        if (ok && tc::elect_one()) {
            const uint32_t idesc = tc::make_idesc_bf16(128, 256);
            const uint64_t da0 = tc::make_smem_desc_sw128(tc::smem_u32(sA), 1024);
            const uint64_t db0 = tc::make_smem_desc_sw128(tc::smem_u32(sW), 1024);
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc::umma_bf16(tmem_base, da0 + (uint64_t)((c * 128 * 128 + k * 32) >> 4),
                                  db0 + (uint64_t)((c * 256 * 128 + k * 32) >> 4), idesc, (c | k) ? 1u : 0u);
            tc::umma_commit(acc_full);
        }
        __syncwarp();
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int m = q * 32 + lane;                       // TMEM lane = gathered row
        const int kp = kp0 + (m >> 2), corner = m & 3;
        float wgt = 0.f;
        bool valid_cell = false;
        if (kp < n_kp) {
            const float x = kp_xy[((size_t)slot * k_cap + kp) * 2 + 0], y = kp_xy[((size_t)slot * k_cap + kp) * 2 + 1];
            const float gx = __fdiv_rn(__fsub_rn(x, 3.5f), (float)img_w - 4.5f);
            const float gy = __fdiv_rn(__fsub_rn(y, 3.5f), (float)img_h - 4.5f);
            const float fx = __fmul_rn(gx, (float)(wc - 1)), fy = __fmul_rn(gy, (float)(hc - 1));
            const float x0f = floorf(fx), y0f = floorf(fy);
            const float ax = __fsub_rn(fx, x0f), ay = __fsub_rn(fy, y0f);
            const int cx = (int)x0f + (corner & 1), cy = (int)y0f + (corner >> 1);
            valid_cell = cx >= 0 && cx < wc && cy >= 0 && cy < hc;
            const float wx = (corner & 1) ? ax : 1.f - ax, wy = (corner >> 1) ? ay : 1.f - ay;
            wgt = __fmul_rn(wx, wy);
        }
        const bool ok = tc::mbar_wait(acc_full, 0, err, 312);
        tc::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        if (ok) {
            // pass 1: norm of this cell's raw 256-d descriptor (F.normalize eps 1e-12)
            float ss = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < 256; c0 += 32) {
                uint32_t r[32];
                tc::tmem_ld32(taddr + c0, r);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) { const float t = __uint_as_float(r[j]) + s_bias[c0 + j]; ss += t * t; }
            }
            const float coef = valid_cell ? wgt * (1.0f / fmaxf(sqrtf(ss), 1e-12f)) : 0.f;
            // pass 2: norm of the interpolated descriptor
            float so = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < 256; c0 += 32) {
                uint32_t r[32];
                tc::tmem_ld32(taddr + c0, r);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float t = (__uint_as_float(r[j]) + s_bias[c0 + j]) * coef;
                    t += __shfl_xor_sync(0xffffffffu, t, 1);
                    t += __shfl_xor_sync(0xffffffffu, t, 2);
                    so += t * t;
                }
            }
            const float inv_o = 1.0f / fmaxf(sqrtf(so), 1e-12f);
            // pass 3: write; the 4 lanes of a keypoint each store 8 of every 32 columns (32 B)
            float* o = desc + ((size_t)slot * k_cap + kp) * 256;
#pragma unroll 1
            for (int c0 = 0; c0 < 256; c0 += 32) {
                uint32_t r[32];
                tc::tmem_ld32(taddr + c0, r);
                tc::tmem_ld_wait();
                float mine[8];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float t = (__uint_as_float(r[j]) + s_bias[c0 + j]) * coef;
                    t += __shfl_xor_sync(0xffffffffu, t, 1);
                    t += __shfl_xor_sync(0xffffffffu, t, 2);
                    if ((j >> 3) == corner) mine[j & 7] = t * inv_o;
                }
                if (kp < n_kp) {
                    float4* dst = reinterpret_cast<float4*>(o + c0 + corner * 8);
                    dst[0] = make_float4(mine[0], mine[1], mine[2], mine[3]);
                    dst[1] = make_float4(mine[4], mine[5], mine[6], mine[7]);
                }
            }
        }
        tc::tc_fence_before();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 256); }
}

int gnb_desc_head_tc(gnb_ctx* ctx, const CUtensorMap* tmap_w, const bf16* ada, const float* bias, int n, int h, int w, int slot0) {
    GNB_CUDA(ctx, gnb_func_smem(ctx, desc_head_kernel, DH_SMEM));
    const int k = ctx->cfg.max_keypoints;
    dim3 grid(ceil_div(k, 32), n);
    GNB_KERNEL(ctx, "desc_head_tc", desc_head_kernel<<<grid, 256, DH_SMEM, ctx->stream>>>(
        *tmap_w, ada, bias, h / 8, w / 8, h, w, ctx->kp_xy, ctx->kp_count, slot0, k, ctx->desc_f32, gnb_tc_err_dev(ctx)));
    return GNB_OK;
}

// ------------------------------------------------------------------------------------------------
// descriptor head on demand, fp32-faithful mode: the gathered cells are [hi: 256 | lo: 256] (eight 16 KB chunks, all
// resident), the split weights [256 rows][hi | lo] are STREAMED through a three-stage TMA ring in the order hi0, lo0, hi1,
// lo1, ... (32 KB per chunk: they do not fit next to the operand), and every weight chunk is multiplied with the
// activation chunks it pairs with (hi c: A_hi[c] and A_lo[c]; lo c: A_hi[c]) into one fp32 accumulator.  Same epilogue as
// desc_head_kernel (per-cell L2 norm, bilinear combine, L2 norm) on the fp32 TMEM values.
#define DX_A_BYTES (8 * 128 * 128)     // 128 KB
#define DX_W_CHUNK (256 * 128)         // 32 KB
#define DX_STAGES 3
#define DX_SMEM (1024 + DX_A_BYTES + DX_STAGES * DX_W_CHUNK + 256 + 256 * 4)

__global__ void __launch_bounds__(256, 1) desc_head_x3_kernel(const __grid_constant__ CUtensorMap tmap_w, const bf16* __restrict__ ada,
                                                              const float* __restrict__ bias, int hc, int wc, int img_h, int img_w,
                                                              const float* __restrict__ kp_xy, const int* __restrict__ kp_count,
                                                              int slot0, int k_cap, float* __restrict__ desc, int* err) {
    const int b = blockIdx.y, slot = slot0 + b;
    const int n_kp = max(kp_count[slot], 0);
    const int kp0 = blockIdx.x * 32;
    if (kp0 >= n_kp) return;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sW = smem + DX_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + DX_STAGES * DX_W_CHUNK);
    uint64_t* w_full = bars;                  // [DX_STAGES]
    uint64_t* w_empty = bars + DX_STAGES;     // [DX_STAGES]
    uint64_t* acc_full = w_empty + DX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < DX_STAGES; ++s) { tc::mbar_init(&w_full[s], 1); tc::mbar_init(&w_empty[s], 1); }
        tc::mbar_init(acc_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) { tc::tmem_alloc(tmem_slot, 256); tc::tmem_relinquish(); }
    s_bias[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    if (warp == 0 && lane == 0) {   // the first ring stages can stream in while the CTA gathers its operand
        for (int q = 0; q < DX_STAGES; ++q) {
            tc::mbar_arrive_expect_tx(&w_full[q], DX_W_CHUNK);
            tc::tma_load_3d(sW + q * DX_W_CHUNK, &tmap_w, &w_full[q], (q & 1) * 256 + (q >> 1) * 64, 0, 0);
        }
    }
    // ---- gather: row m = 4 * (keypoint - kp0) + corner; 1024 B per row = 8 chunks x 8 sixteen-byte pieces (hi chunks 0-3, lo 4-7)
    const bf16* base = ada + (size_t)b * hc * wc * 512;
    for (int item = threadIdx.x; item < 128 * 64; item += 256) {
        const int m = item >> 6, piece = item & 63;
        const int kp = kp0 + (m >> 2), corner = m & 3;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (kp < n_kp) {
            const float x = kp_xy[((size_t)slot * k_cap + kp) * 2 + 0], y = kp_xy[((size_t)slot * k_cap + kp) * 2 + 1];
            const float gx = __fdiv_rn(__fsub_rn(x, 3.5f), (float)img_w - 4.5f);
            const float gy = __fdiv_rn(__fsub_rn(y, 3.5f), (float)img_h - 4.5f);
            const int cx = (int)floorf(__fmul_rn(gx, (float)(wc - 1))) + (corner & 1);
            const int cy = (int)floorf(__fmul_rn(gy, (float)(hc - 1))) + (corner >> 1);
            if (cx >= 0 && cx < wc && cy >= 0 && cy < hc)
                val = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)cy * wc + cx) * 512) + piece);
        }
        const int chunk = piece >> 3, j = piece & 7;
        *reinterpret_cast<uint4*>(sA + chunk * 128 * 128 + m * 128 + ((j ^ (m & 7)) << 4)) = val;
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int q = DX_STAGES; q < 8; ++q) {
                const int s = q % DX_STAGES;
                if (!tc::mbar_wait(&w_empty[s], ((q / DX_STAGES) & 1) ^ 1, err, 331)) break;
                tc::mbar_arrive_expect_tx(&w_full[s], DX_W_CHUNK);
                tc::tma_load_3d(sW + s * DX_W_CHUNK, &tmap_w, &w_full[s], (q & 1) * 256 + (q >> 1) * 64, 0, 0);
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = tc::make_idesc_bf16(128, 256);
        const uint64_t da0 = tc::make_smem_desc_sw128(tc::smem_u32(sA), 1024);
        const uint64_t db0 = tc::make_smem_desc_sw128(tc::smem_u32(sW), 1024);
        bool ok = true;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int s = q % DX_STAGES;
            if (ok && !tc::mbar_wait(&w_full[s], (q / DX_STAGES) & 1, err, 332)) ok = false;
            tc::tc_fence_after();
            if (ok && tc::elect_one()) {
                const int c = q >> 1;
                const uint64_t db = db0 + (uint64_t)((s * DX_W_CHUNK) >> 4);
#pragma unroll
                for (int term = 0; term < 2; ++term) {
                    if (term < ((q & 1) ? 1 : 2)) {   // hi weights: A_hi[c], A_lo[c];  lo weights: A_hi[c]
                        const uint64_t da = da0 + (uint64_t)(((c + term * 4) * 128 * 128) >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k) tc::umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (q | term | k) ? 1u : 0u);
                    }
                }
                tc::umma_commit(&w_empty[s]);
                if (q == 7) tc::umma_commit(acc_full);
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int m = q * 32 + lane;                       // TMEM lane = gathered row
        const int kp = kp0 + (m >> 2), corner = m & 3;
        float wgt = 0.f;
        bool valid_cell = false;
        if (kp < n_kp) {
            const float x = kp_xy[((size_t)slot * k_cap + kp) * 2 + 0], y = kp_xy[((size_t)slot * k_cap + kp) * 2 + 1];
            const float gx = __fdiv_rn(__fsub_rn(x, 3.5f), (float)img_w - 4.5f);
            const float gy = __fdiv_rn(__fsub_rn(y, 3.5f), (float)img_h - 4.5f);
            const float fx = __fmul_rn(gx, (float)(wc - 1)), fy = __fmul_rn(gy, (float)(hc - 1));
            const float x0f = floorf(fx), y0f = floorf(fy);
            const float ax = __fsub_rn(fx, x0f), ay = __fsub_rn(fy, y0f);
            const int cx = (int)x0f + (corner & 1), cy = (int)y0f + (corner >> 1);
            valid_cell = cx >= 0 && cx < wc && cy >= 0 && cy < hc;
            const float wx = (corner & 1) ? ax : 1.f - ax, wy = (corner >> 1) ? ay : 1.f - ay;
            wgt = __fmul_rn(wx, wy);
        }
        const bool ok = tc::mbar_wait(acc_full, 0, err, 333);
        tc::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        if (ok) {
            float ss = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < 256; c0 += 32) {
                uint32_t r[32];
                tc::tmem_ld32(taddr + c0, r);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) { const float t = __uint_as_float(r[j]) + s_bias[c0 + j]; ss += t * t; }
            }
            const float coef = valid_cell ? wgt * (1.0f / fmaxf(sqrtf(ss), 1e-12f)) : 0.f;
            float so = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < 256; c0 += 32) {
                uint32_t r[32];
                tc::tmem_ld32(taddr + c0, r);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float t = (__uint_as_float(r[j]) + s_bias[c0 + j]) * coef;
                    t += __shfl_xor_sync(0xffffffffu, t, 1);
                    t += __shfl_xor_sync(0xffffffffu, t, 2);
                    so += t * t;
                }
            }
            const float inv_o = 1.0f / fmaxf(sqrtf(so), 1e-12f);
            float* o = desc + ((size_t)slot * k_cap + kp) * 256;
#pragma unroll 1
            for (int c0 = 0; c0 < 256; c0 += 32) {
                uint32_t r[32];
                tc::tmem_ld32(taddr + c0, r);
                tc::tmem_ld_wait();
                float mine[8];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float t = (__uint_as_float(r[j]) + s_bias[c0 + j]) * coef;
                    t += __shfl_xor_sync(0xffffffffu, t, 1);
                    t += __shfl_xor_sync(0xffffffffu, t, 2);
                    if ((j >> 3) == corner) mine[j & 7] = t * inv_o;
                }
                if (kp < n_kp) {
                    float4* dst = reinterpret_cast<float4*>(o + c0 + corner * 8);
                    dst[0] = make_float4(mine[0], mine[1], mine[2], mine[3]);
                    dst[1] = make_float4(mine[4], mine[5], mine[6], mine[7]);
                }
            }
        }
        tc::tc_fence_before();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 256); }
}

int gnb_desc_head_x3_tc(gnb_ctx* ctx, const CUtensorMap* tmap_w, const bf16* ada, const float* bias, int n, int h, int w, int slot0) {
    GNB_CUDA(ctx, gnb_func_smem(ctx, desc_head_x3_kernel, DX_SMEM));
    const int k = ctx->cfg.max_keypoints;
    dim3 grid(ceil_div(k, 32), n);
    GNB_KERNEL(ctx, "desc_head_x3", desc_head_x3_kernel<<<grid, 256, DX_SMEM, ctx->stream>>>(
        *tmap_w, ada, bias, h / 8, w / 8, h, w, ctx->kp_xy, ctx->kp_count, slot0, k, ctx->desc_f32, gnb_tc_err_dev(ctx)));
    return GNB_OK;
}

// ------------------------------------------------------------------------------------------------
// LightGlue final projection on tcgen05: m = (W d + b) / 256^(1/4) for 128 keypoints per CTA, plus
// the matchability logit z = w_m . d + b_m as a second, 16-column MMA on the same A operand
// (B2 = [w_m; 0 ...]).  A = bf16(descriptors) gathered row-wise into the 128B-swizzle layout.
#define PJ_W2_BYTES (4 * 16 * 128)   // four K chunks of [16 rows x 128 B]; row 0 = w_m
#define PJ_SMEM (1024 + DH_W_BYTES + PJ_W2_BYTES + DH_A_BYTES + 256 + 256 * 4)

__global__ void __launch_bounds__(256, 1) project_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const float* __restrict__ desc,
                                                            const int* __restrict__ kp_count, int slot0, int k_cap,
                                                            const float* __restrict__ bias, const bf16* __restrict__ mw, float mb,
                                                            bf16* __restrict__ mproj, float* __restrict__ mlogit, int* err) {
    const int slot = slot0 + blockIdx.y;
    const int n_kp = max(kp_count[slot], 0);
    const int kp0 = blockIdx.x * 128;
    if (kp0 >= n_kp) return;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW = smem;
    uint8_t* sW2 = sW + DH_W_BYTES;
    uint8_t* sA = sW2 + PJ_W2_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + DH_A_BYTES);
    uint64_t* w_full = bars;
    uint64_t* acc_full = bars + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmap_w);
        tc::mbar_init(w_full, 1);
        tc::mbar_init(acc_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) { tc::tmem_alloc(tmem_slot, 512); tc::tmem_relinquish(); }
    s_bias[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    if (warp == 0 && lane == 0) {
        tc::mbar_arrive_expect_tx(w_full, DH_W_BYTES);
        for (int c = 0; c < 4; ++c) tc::tma_load_3d(sW + c * 256 * 128, &tmap_w, w_full, c * 64, 0, 0);
    }
    // B2: row 0 = w_m (bf16), rows 1..15 = 0
    for (int item = threadIdx.x; item < 4 * 16 * 8; item += 256) {
        const int chunk = item >> 7, row = (item >> 3) & 15, j = item & 7;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (row == 0) val = *reinterpret_cast<const uint4*>(mw + chunk * 64 + j * 8);
        *reinterpret_cast<uint4*>(sW2 + chunk * 16 * 128 + row * 128 + ((j ^ (row & 7)) << 4)) = val;
    }
    // A: 128 keypoints x 256 ch, f32 -> bf16; item = (row, group of 8 channels)
    const float* dbase = desc + ((size_t)slot * k_cap + kp0) * 256;
    for (int item = threadIdx.x; item < 128 * 32; item += 256) {
        const int m = item >> 5, piece = item & 31;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (kp0 + m < n_kp) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(dbase + (size_t)m * 256 + piece * 8));
            const float4 b = __ldg(reinterpret_cast<const float4*>(dbase + (size_t)m * 256 + piece * 8 + 4));
            val = make_uint4(tc::pack_bf16x2(a.x, a.y), tc::pack_bf16x2(a.z, a.w), tc::pack_bf16x2(b.x, b.y), tc::pack_bf16x2(b.z, b.w));
        }
        const int chunk = piece >> 3, j = piece & 7;
        *reinterpret_cast<uint4*>(sA + chunk * 128 * 128 + m * 128 + ((j ^ (m & 7)) << 4)) = val;
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 1) {
        const bool ok = tc::mbar_wait(w_full, 0, err, 321);
        if (ok && tc::elect_one()) {
            const uint32_t idesc = tc::make_idesc_bf16(128, 256), idesc2 = tc::make_idesc_bf16(128, 16);
            const uint64_t da0 = tc::make_smem_desc_sw128(tc::smem_u32(sA), 1024);
            const uint64_t db0 = tc::make_smem_desc_sw128(tc::smem_u32(sW), 1024);
            const uint64_t db2 = tc::make_smem_desc_sw128(tc::smem_u32(sW2), 1024);
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t da = da0 + (uint64_t)((c * 128 * 128 + k * 32) >> 4);
                    tc::umma_bf16(tmem_base, da, db0 + (uint64_t)((c * 256 * 128 + k * 32) >> 4), idesc, (c | k) ? 1u : 0u);
                    tc::umma_bf16(tmem_base + 256u, da, db2 + (uint64_t)((c * 16 * 128 + k * 32) >> 4), idesc2, (c | k) ? 1u : 0u);
                }
            tc::umma_commit(acc_full);
        }
        __syncwarp();
    } else if (warp >= 4) {
        const int q = warp & 3, m = q * 32 + lane, kp = kp0 + m;
        const bool ok = tc::mbar_wait(acc_full, 0, err, 322);
        tc::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        if (ok) {
            bf16* o = mproj + ((size_t)slot * k_cap + kp) * 256;
#pragma unroll 1
            for (int c0 = 0; c0 < 256; c0 += 32) {
                uint32_t r[32];
                tc::tmem_ld32(taddr + c0, r);
                tc::tmem_ld_wait();
                if (kp < n_kp) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint32_t pk[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int c = c0 + g * 8 + 2 * j;
                            pk[j] = tc::pack_bf16x2(__fmul_rn(__fadd_rn(__uint_as_float(r[g * 8 + 2 * j]), s_bias[c]), 0.25f),
                                                    __fmul_rn(__fadd_rn(__uint_as_float(r[g * 8 + 2 * j + 1]), s_bias[c + 1]), 0.25f));
                        }
                        *reinterpret_cast<uint4*>(o + c0 + g * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
            }
            uint32_t r[32];
            tc::tmem_ld32(taddr + 256u, r);
            tc::tmem_ld_wait();
            if (kp < n_kp) {
                const float z = __uint_as_float(r[0]) + mb;
                mlogit[(size_t)slot * k_cap + kp] = fminf(z, 0.f) - log1pf(expf(-fabsf(z)));
            }
        }
        tc::tc_fence_before();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}

int gnb_project_tc(gnb_ctx* ctx, int slot0, int n_slots) {
    TcState* ts = tc_state(ctx);
    CUtensorMap& g_proj_wmap = ts->proj_map;
    if (!ts->proj_ready) {
        const uint64_t dims[3] = {256, 256, 1};
        const uint64_t strides[2] = {512, 256 * 512};
        const uint32_t box[3] = {64, 256, 1};
        int rc = gnb_make_tmap_bf16(ctx, &g_proj_wmap, ctx->match_w, 3, dims, strides, box);
        if (rc) return rc;
        GNB_CUDA(ctx, cudaFuncSetAttribute(project_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PJ_SMEM));
        ts->proj_ready = 1;
    }
    const int k = ctx->cfg.max_keypoints;
    dim3 grid(ceil_div(k, 128), n_slots);
    GNB_KERNEL(ctx, "project_tc", project_tc_kernel<<<grid, 256, PJ_SMEM, ctx->stream>>>(
        g_proj_wmap, ctx->desc_f32, ctx->kp_count, slot0, k, ctx->match_b, ctx->match_mw, ctx->match_mb, ctx->mproj, ctx->mlogit,
        gnb_tc_err_dev(ctx)));
    return GNB_OK;
}
