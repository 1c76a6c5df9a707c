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

#include <stdlib.h>

#define CT_TH 8
#define CT_TW 16
#define CT_A_BYTES (128 * 128)   // 128 pixels x 64 ch x 2 B
#define CT_STAGES 4

int* gnb_tc_err_dev(gnb_ctx* ctx);


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
        const uint32_t idesc = tc::make_idesc_bf16(128, NPAD);
        const uint64_t d0 = tc::make_smem_desc_sw128(tc::smem_u32(smem), 1024);
        for (int it = 0; it < n_iters; ++it) {
            const int s = it % CT_STAGES;
            const uint32_t ph = (it / CT_STAGES) & 1;
            if (!tc::mbar_wait(&full[s], ph, err, 202)) break;
            tc::tc_fence_after();
            if (tc::elect_one()) {
                const uint64_t da0 = d0 + (uint64_t)((s * STAGE_BYTES) >> 4);
                const uint64_t db0 = da0 + (uint64_t)(CT_A_BYTES >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) tc::umma_bf16(tmem_base, da0 + 2 * k, db0 + 2 * k, idesc, (it | k) ? 1u : 0u);
                tc::umma_commit(&empty[s]);
            }
            __syncwarp();
        }
        if (tc::elect_one()) tc::umma_commit(acc_full);
        __syncwarp();
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

// ------------------------------------------------------------------------------------------------
// v2 for Cin = 64 layers (conv1b, conv2a, conv2b, conv3a: 70 % of the stack's FLOPs):
//   * persistent CTAs (one per SM) looping over 16 x 8 output tiles;
//   * the layer's whole weight tensor (9 taps x Cout x 128 B) stays resident in shared memory;
//   * ONE TMA box per tile: the 18 x 10 halo (64 ch = 128 B per pixel).  The nine tap operands are
//     nine views of that box: the UMMA descriptor's start address is shifted by (dy*10 + dx) pixels
//     (128 B each) and its stride-byte-offset is the halo row pitch (10 x 128 B), so row group g of
//     the M = 128 operand is image row g of the tile.  The 128B swizzle is a function of the shared
//     memory address, so TMA's write pattern and the shifted reads agree;
//   * two TMEM accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
// L2 -> SM traffic drops from 9 x (16 KB + weights) to 22.5 KB per tile.
#define H2_TH 16
#define H2_TW 8
#define H2_HH (H2_TH + 2)
#define H2_HW (H2_TW + 2)
#define H2_HALO_BYTES (H2_HH * H2_HW * 128)           // 23040
#define H2_HALO_STRIDE ((H2_HALO_BYTES + 1023) / 1024 * 1024)

// KCH = Cin / 64 (1 or 2).  With KCH = 2 the A ring holds one (tile, 64-channel chunk) halo box per
// stage and the output channels are split into `n_slices` slices of NPAD channels: a CTA keeps the
// weights of ONE slice resident for the whole layer (blockIdx.x % n_slices) and loops over tiles,
// so the only streamed operand is the 22.5 KB halo box (L2 traffic per tile = n_slices x KCH x 22.5 KB).
//
// X3 = the fp32-faithful mode (cfg.precision = 1, oracle/superpoint_ref.py quantize="x3"): activations and weights
// travel as TWO bf16 terms per value, v = hi + lo, stored as channel blocks [hi: C | lo: C] of one NHWC tensor, so
// the KCH = 2 C / 64 chunks of a pixel are the hi chunks followed by the lo chunks.  The product keeps
// a_hi w_hi + a_hi w_lo + a_lo w_hi: a hi chunk is multiplied with the hi AND the lo weight block of the same
// channels, a lo chunk with the hi block only — three MMAs into the same fp32 TMEM accumulator per (tap, k-step).
// The epilogue splits the fp32 result again (hi = bf16(v), lo = bf16(v - hi)); pooling is done on the fp32 values.
template <int NPAD, int KCH, int STAGES, bool X3 = false>
__global__ void __launch_bounds__(256, 1) conv_tc_halo_kernel(const __grid_constant__ CUtensorMap tmap_in,
                                                              const __grid_constant__ CUtensorMap tmap_w,
                                                              const float* __restrict__ bias, int h, int w, int n_img, int cout,
                                                              int n_slices, bf16* __restrict__ out_bf, int relu, int pool, int* err,
                                                              const __grid_constant__ CUtensorMap tmap_out, int tstore) {
    constexpr int W_TAP_BYTES = NPAD * 128;        // one (tap, chunk) block
    constexpr int W_BYTES = 9 * KCH * W_TAP_BYTES;
    // X3: the hi chunk is multiplied with the hi AND lo weight blocks in ONE MMA of N = 2 NPAD (the two blocks of a tap are
    // contiguous in shared memory: 8 KB of operands per 64-cycle instruction instead of 2 x 6 KB per 2 x 32 cycles), which
    // leaves two partial sums per channel — columns [0, NPAD): a_hi w_hi (+ a_lo w_hi from the lo chunk), [NPAD, 2 NPAD):
    // a_hi w_lo — that the epilogue adds.
    constexpr int ACC_COLS = X3 ? 2 * NPAD : NPAD;
    constexpr int TMEM_COLS = 2 * ACC_COLS;  // 128 or 256
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW = smem;
    uint8_t* sA = smem + W_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + STAGES * H2_HALO_STRIDE);
    uint64_t* w_full = bars;
    uint64_t* a_full = bars + 1;             // [STAGES]
    uint64_t* a_empty = a_full + STAGES;     // [STAGES]
    uint64_t* t_full = a_empty + STAGES;     // [2]
    uint64_t* t_empty = t_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [NPAD], 16-byte aligned
    // tstore: 16 KB staging panel [128 px][64 ch] (128B swizzle) for the TMA write-out, 1024-byte aligned behind the bias
    uint8_t* sOut = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_bias + NPAD) + 1023) & ~uintptr_t(1023));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (w + H2_TW - 1) / H2_TW, tiles_y = (h + H2_TH - 1) / H2_TH;
    const int tiles_per_img = tiles_x * tiles_y;
    const int total = tiles_per_img * n_img;
    const int slice = blockIdx.x % n_slices;           // output-channel slice this CTA owns
    const int tile0 = blockIdx.x / n_slices;           // first tile, then stride = CTAs per slice
    const int tstride = gridDim.x / n_slices;
    const int ch0 = slice * NPAD;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmap_in);
        tc::tma_prefetch_desc(&tmap_w);
        tc::mbar_init(w_full, 1);
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&a_full[s], 1); tc::mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&t_full[s], 1); tc::mbar_init(&t_empty[s], 4); }
        tc::fence_barrier_init();
    }
    if (warp == 2) {
        tc::tmem_alloc(tmem_slot, TMEM_COLS);
        tc::tmem_relinquish();
    }
    if (threadIdx.x >= 128 && threadIdx.x - 128 < NPAD) s_bias[threadIdx.x - 128] = bias[ch0 + threadIdx.x - 128];
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(w_full, W_BYTES);
            for (int t = 0; t < 9; ++t)
                for (int c = 0; c < KCH; ++c)
                    tc::tma_load_3d(sW + (t * KCH + c) * W_TAP_BYTES, &tmap_w, w_full, c * 64, ch0, t);
            int i = 0;  // ring index over (tile, chunk) boxes
            for (int tile = tile0; tile < total; tile += tstride) {
                const int img = tile / tiles_per_img, rem = tile % tiles_per_img;
                const int y0 = (rem / tiles_x) * H2_TH, x0 = (rem % tiles_x) * H2_TW;
                bool ok = true;
                for (int c = 0; c < KCH; ++c, ++i) {
                    const int s = i % STAGES;
                    const uint32_t ph = (i / STAGES) & 1;
                    if (i >= STAGES && !tc::mbar_wait(&a_empty[s], ph ^ 1, err, 211)) { ok = false; break; }
                    tc::mbar_arrive_expect_tx(&a_full[s], H2_HALO_BYTES);
                    tc::tma_load_4d(sA + s * H2_HALO_STRIDE, &tmap_in, &a_full[s], c * 64, x0 - 1, y0 - 1, img);
                }
                if (!ok) break;
            }
        }
    } else if (warp == 1) {
        // whole warp stays converged; one elected lane issues.  Descriptors are a 64-bit base plus a
        // compile-time constant per (tap, k-step): one add each.
        const uint32_t idesc = tc::make_idesc_bf16(128, NPAD);
        const uint32_t idesc2 = tc::make_idesc_bf16(128, X3 ? 2 * NPAD : NPAD);   // X3 hi chunk: N covers the hi and the lo weight block
        bool ok = tc::mbar_wait(w_full, 0, err, 212);
        const uint64_t db0 = tc::make_smem_desc_sw128(tc::smem_u32(sW), 1024);
        int i = 0, ti = 0;
        for (int tile = tile0; ok && tile < total; tile += tstride, ++ti) {
            const int as = ti & 1;
            if (ti >= 2 && !tc::mbar_wait(&t_empty[as], ((ti >> 1) & 1) ^ 1, err, 214)) break;
            const uint32_t d_tmem = tmem_base + (uint32_t)(as * ACC_COLS);
#pragma unroll
            for (int c = 0; c < KCH; ++c, ++i) {
                const int s = i % STAGES;
                if (!tc::mbar_wait(&a_full[s], (i / STAGES) & 1, err, 213)) { ok = false; break; }
                tc::tc_fence_after();
                const uint64_t da0 = tc::make_smem_desc_sw128(tc::smem_u32(sA + s * H2_HALO_STRIDE), H2_HW * 128);
                if (tc::elect_one()) {
                    // X3 (KCH = 2): chunk 0 = hi activations x [W_hi; W_lo] (one N = 2 NPAD MMA per tap and k-step), chunk 1 = lo
                    // activations x W_hi (N = NPAD, accumulating into the first NPAD columns)
                    static_assert(!X3 || KCH == 2, "the split-bf16 single-CTA kernel is written for Cin = 64");
                    const int wc = X3 ? 0 : c;
                    const uint32_t id = (X3 && c == 0) ? idesc2 : idesc;
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
#pragma unroll
                        for (int k2 = 0; k2 < 4; ++k2) {
                            const uint64_t da = da0 + (uint64_t)((((t / 3) * H2_HW + (t % 3)) * 128 + k2 * 32) >> 4);
                            const uint64_t db = db0 + (uint64_t)(((t * KCH + wc) * W_TAP_BYTES + k2 * 32) >> 4);
                            tc::umma_bf16(d_tmem, da, db, id, (c | t | k2) ? 1u : 0u);
                        }
                    }
                    tc::umma_commit(&a_empty[s]);
                    if (c == KCH - 1) tc::umma_commit(&t_full[as]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int yl = 4 * q + (lane >> 3), xl = lane & 7;
        const int opitch = X3 ? 2 * cout : cout;     // output channels per pixel (X3: hi block then lo block)
        int i = 0;
        for (int tile = tile0; tile < total; tile += tstride, ++i) {
            const int as = i & 1;
            const int img = tile / tiles_per_img, rem = tile % tiles_per_img;
            const int y = (rem / tiles_x) * H2_TH + yl, x = (rem % tiles_x) * H2_TW + xl;
            if (!tc::mbar_wait(&t_full[as], (i >> 1) & 1, err, 215)) break;
            tc::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * ACC_COLS);
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
#pragma unroll 1
            for (int c0 = 0; c0 < NPAD; c0 += 32) {
                uint32_t v[32];
                tc::tmem_ld32(taddr + c0, v);
                if constexpr (X3) {   // add the a_hi w_lo partial sums
                    uint32_t v2[32];
                    tc::tmem_ld32(taddr + NPAD + c0, v2);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__fadd_rn(__uint_as_float(v[j]), __uint_as_float(v2[j])));
                }
                tc::tmem_ld_wait();
                if constexpr (X3) {
                    uint32_t phi[16], plo[16];
                    tc::epilogue_split32(v, &s_bias[c0], relu, pool, phi, plo);
                    if (writer) {
                        uint4* oh = reinterpret_cast<uint4*>(out_bf + pix * opitch + ch0 + c0);
                        uint4* ol = reinterpret_cast<uint4*>(out_bf + pix * opitch + cout + ch0 + c0);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            oh[g] = make_uint4(phi[4 * g], phi[4 * g + 1], phi[4 * g + 2], phi[4 * g + 3]);
                            ol[g] = make_uint4(plo[4 * g], plo[4 * g + 1], plo[4 * g + 2], plo[4 * g + 3]);
                        }
                    }
                } else {
                uint32_t packed[16];
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 bb = *reinterpret_cast<const float4*>(&s_bias[c0 + 4 * j4]);
                    const float a0 = __uint_as_float(v[4 * j4]) + bb.x, a1 = __uint_as_float(v[4 * j4 + 1]) + bb.y;
                    const float a2 = __uint_as_float(v[4 * j4 + 2]) + bb.z, a3 = __uint_as_float(v[4 * j4 + 3]) + bb.w;
                    packed[2 * j4] = relu ? tc::pack_bf16x2_relu(a0, a1) : tc::pack_bf16x2(a0, a1);
                    packed[2 * j4 + 1] = relu ? tc::pack_bf16x2_relu(a2, a3) : tc::pack_bf16x2(a2, a3);
                }
                if (pool) {  // max of bf16-rounded values == rounding of the max (monotone)
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        uint32_t u = packed[j];
                        uint32_t u1 = __shfl_xor_sync(0xffffffffu, u, 1);
                        __nv_bfloat162 o = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&u), *reinterpret_cast<__nv_bfloat162*>(&u1));
                        u = *reinterpret_cast<uint32_t*>(&o);
                        uint32_t u8 = __shfl_xor_sync(0xffffffffu, u, 8);
                        o = __hmax2(o, *reinterpret_cast<__nv_bfloat162*>(&u8));
                        packed[j] = *reinterpret_cast<uint32_t*>(&o);
                    }
                }
                if (tstore) {
                    // full-resolution write-out by TMA: the tile's 64-channel panel is staged as [128 px][128 B] (pixel m =
                    // yl * 8 + xl = this thread's TMEM lane) and stored with one bulk tensor copy: whole lines, clipped at
                    // the image border by the tensor map, and no store instruction competes with the MMA operand fetch
                    const int m = q * 32 + lane;
                    if ((c0 & 63) == 0) {
                        if (threadIdx.x == 128) tc::tma_store_wait_read();   // previous panel has left the staging buffer
                        tc::named_bar_sync(1, 128);
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        *reinterpret_cast<uint4*>(sOut + m * 128 + (((((c0 & 63) >> 3) + g) ^ (m & 7)) << 4)) =
                            make_uint4(packed[4 * g], packed[4 * g + 1], packed[4 * g + 2], packed[4 * g + 3]);
                    if ((c0 & 63) == 32) {
                        tc::fence_proxy_async_smem();
                        tc::named_bar_sync(1, 128);
                        if (threadIdx.x == 128) {
                            tc::tma_store_4d(&tmap_out, sOut, ch0 + c0 - 32, x - xl, y - yl, img);
                            tc::tma_store_commit();
                        }
                    }
                } else if (writer) {
                    uint4* o = reinterpret_cast<uint4*>(out_bf + pix * cout + ch0 + c0);
#pragma unroll
                    for (int g = 0; g < 4; ++g) o[g] = make_uint4(packed[4 * g], packed[4 * g + 1], packed[4 * g + 2], packed[4 * g + 3]);
                }
                }   // !X3
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&t_empty[as]);
        }
        if (tstore && threadIdx.x == 128) tc::tma_store_wait_all();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int NPAD, int KCH, int STAGES, bool TSTORE = false, bool X3 = false>
static int launch_conv_tc_halo(gnb_ctx* ctx, const CUtensorMap& tin, const CUtensorMap& tw, const ConvLayer& L, int n, int h, int w,
                               bf16* out_bf, int relu, int pool, const char* name) {
    constexpr int smem = 1024 + 9 * KCH * NPAD * 128 + STAGES * H2_HALO_STRIDE + 256 + NPAD * 4 + (TSTORE ? 1024 + 16384 : 0);
    GNB_CUDA(ctx, gnb_func_smem(ctx, conv_tc_halo_kernel<NPAD, KCH, STAGES, X3>, smem));
    const int n_slices = L.cout_pad / NPAD;
    const int total = ceil_div(w, H2_TW) * ceil_div(h, H2_TH) * n;
    int per_slice = ctx->sm_count / n_slices;
    if (per_slice > total) per_slice = total;
    if (per_slice < 1) per_slice = 1;
    const int grid = per_slice * n_slices;
    CUtensorMap tout = tin;   // placeholder when the TMA write-out is off
    const int tstore = (TSTORE && !pool) ? 1 : 0;
    if (tstore) {
        const uint64_t od[4] = {(uint64_t)L.cout, (uint64_t)w, (uint64_t)h, (uint64_t)n};
        const uint64_t os[3] = {(uint64_t)L.cout * 2, (uint64_t)w * L.cout * 2, (uint64_t)h * w * L.cout * 2};
        const uint32_t ob[4] = {64, H2_TW, H2_TH, 1};
        int rc = gnb_make_tmap_bf16(ctx, &tout, out_bf, 4, od, os, ob);
        if (rc) return rc;
    }
    GNB_KERNEL(ctx, name, conv_tc_halo_kernel<NPAD, KCH, STAGES, X3><<<grid, 256, smem, ctx->stream>>>(
        tin, tw, L.bias, h, w, n, L.cout, n_slices, out_bf, relu, pool, gnb_tc_err_dev(ctx), tout, tstore));
    return GNB_OK;
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant for the Cin = 128 layers (conv3b, conv4a, conv4b, convPa, convDa).
// A single SM cannot hold the weights of 128 output channels (9 x 2 x 128 x 128 B = 288 KB), so the kernel above
// runs N = 64 slices, and N = 64 MMAs are bound by shared-memory operand bandwidth (6 KB per 32-cycle instruction).
// Here two SMs of a TPC form a cluster and issue ONE tcgen05.mma.cta_group::2 of M = 256, N = 128 per (tap, k-step):
// each CTA contributes its own 128-pixel halo tile as A and HALF of the 128 weight rows as B (144 KB resident per CTA,
// as before), and receives its 128 pixels x 128 channels in its own TMEM.  Operand traffic per SM drops to 6 KB per
// 64-cycle instruction (96 B/clk): the MMA runs at the tensor rate instead of the shared-memory rate.
//   leader (cluster rank 0): MMA issuer.  It waits for its own and the peer's halo box (the peer relays its TMA
//   completion with a remote mbarrier arrive), and its tcgen05.commit is multicast to the barriers of both CTAs.
//   both: TMA producer for their own tile / weight half, epilogue for their own accumulator; the peer's epilogue
//   warps release the accumulator stage on the leader's barrier.
// NP = output channels per pair-slice (128; 64 in the fp32-faithful X3 mode, whose 2x as many resident weight blocks
// leave room for 32 weight rows per CTA), X3 as in conv_tc_halo_kernel.
template <int KCH, int STAGES, int NP = 128, bool X3 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
conv_tc_halo_pair_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_w,
                         const float* __restrict__ bias, int h, int w, int n_img, int cout, int n_slices,
                         bf16* __restrict__ out_bf, int relu, int pool, int* err) {
    constexpr int HP_N = NP, HP_NH = NP / 2;
    constexpr int W_TAP_BYTES = HP_NH * 128;
    constexpr int W_BYTES = 9 * KCH * W_TAP_BYTES;
    // X3 (KCH = 4: hi0, hi1, lo0, lo1): a hi chunk is multiplied with its hi AND lo weight rows in ONE pair MMA of
    // N = 2 NP (each CTA keeps the hi and lo rows of its NP / 2 channels contiguous: blocks are stored hi0, lo0, hi1, lo1),
    // a lo chunk with the hi rows (N = NP) into a second accumulator region: three partial sums per channel, added by the
    // epilogue.  Columns of a stage: [0, 2 NP): per CTA half [hh | hl], [2 NP, 3 NP): lh.
    constexpr int ACC_COLS = X3 ? 4 * HP_N : HP_N;
    constexpr int TMEM_COLS = 2 * ACC_COLS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW = smem;
    uint8_t* sA = smem + W_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + STAGES * H2_HALO_STRIDE);
    uint64_t* w_full = bars;
    uint64_t* a_full = bars + 1;             // [STAGES] own halo box landed
    uint64_t* a_empty = a_full + STAGES;     // [STAGES] MMAs done reading (multicast commit)
    uint64_t* t_full = a_empty + STAGES;     // [2] accumulator ready (multicast commit)
    uint64_t* t_empty = t_full + 2;          // [2] leader only: 4 local + 4 remote epilogue warps
    uint64_t* pa_full = t_empty + 2;         // [STAGES] leader only: the peer's halo box landed
    uint64_t* pw_full = pa_full + STAGES;    // leader only: the peer's weights landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pw_full + 1);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [HP_N]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const int tiles_x = (w + H2_TW - 1) / H2_TW, tiles_y = (h + H2_TH - 1) / H2_TH;
    const int tiles_per_img = tiles_x * tiles_y;
    const int total = tiles_per_img * n_img;
    const int tile_pairs = (total + 1) / 2;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int slice = pair % n_slices;
    const int tp0 = pair / n_slices, tpstride = n_pairs / n_slices;
    const int ch0 = slice * HP_N;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmap_in);
        tc::tma_prefetch_desc(&tmap_w);
        tc::mbar_init(w_full, 1);
        tc::mbar_init(pw_full, 1);
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&a_full[s], 1); tc::mbar_init(&a_empty[s], 1); tc::mbar_init(&pa_full[s], 1); }
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&t_full[s], 1); tc::mbar_init(&t_empty[s], 8); }
        tc::fence_barrier_init();
    }
    if (warp == 2) {
        tc::tmem_alloc_pair(tmem_slot, TMEM_COLS);
        tc::tmem_relinquish_pair();
    }
    if (threadIdx.x >= 128 && threadIdx.x - 128 < HP_N) s_bias[threadIdx.x - 128] = bias[ch0 + threadIdx.x - 128];
    tc::tc_fence_before();
    __syncthreads();
    tc::cluster_sync_all();     // the peer's barriers exist before anything arrives on them remotely
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(w_full, W_BYTES);
            for (int t = 0; t < 9; ++t)
                for (int c = 0; c < KCH; ++c)
                    tc::tma_load_3d(sW + (t * KCH + (X3 ? (c < KCH / 2 ? 2 * c : 2 * (c - KCH / 2) + 1) : c)) * W_TAP_BYTES, &tmap_w, w_full, c * 64,
                                    ch0 + (int)rank * HP_NH, t);
            int i = 0;
            for (int tp = tp0; tp < tile_pairs; tp += tpstride) {
                const int tile = 2 * tp + (int)rank;     // may be == total for the odd tail: the box is then all zero fill
                const int img = tile / tiles_per_img, rem = tile % tiles_per_img;
                const int y0 = (rem / tiles_x) * H2_TH, x0 = (rem % tiles_x) * H2_TW;
                bool ok = true;
                for (int c = 0; c < KCH; ++c, ++i) {
                    const int s = i % STAGES;
                    const uint32_t ph = (i / STAGES) & 1;
                    if (i >= STAGES && !tc::mbar_wait_cluster(&a_empty[s], ph ^ 1, err, 231)) { ok = false; break; }
                    tc::mbar_arrive_expect_tx(&a_full[s], H2_HALO_BYTES);
                    tc::tma_load_4d(sA + s * H2_HALO_STRIDE, &tmap_in, &a_full[s], c * 64, x0 - 1, y0 - 1, img);
                }
                if (!ok) break;
            }
        }
    } else if (warp == 3 && rank == 1) {
        // peer relay: tell the leader when this CTA's operands have landed
        if (lane == 0) {
            if (tc::mbar_wait(w_full, 0, err, 232)) tc::mbar_arrive_cluster(tc::map_to_cta(pw_full, 0));
            int i = 0;
            bool ok = true;
            for (int tp = tp0; ok && tp < tile_pairs; tp += tpstride) {
                for (int c = 0; c < KCH; ++c, ++i) {
                    const int s = i % STAGES;
                    if (!tc::mbar_wait(&a_full[s], (i / STAGES) & 1, err, 233)) { ok = false; break; }
                    tc::mbar_arrive_cluster(tc::map_to_cta(&pa_full[s], 0));
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // leader: one elected lane issues the pair-wide MMAs
        const uint32_t idesc = tc::make_idesc_bf16(256, HP_N);
        const uint32_t idesc2 = tc::make_idesc_bf16(256, X3 ? 2 * HP_N : HP_N);
        bool ok = tc::mbar_wait(w_full, 0, err, 234) && tc::mbar_wait_cluster(pw_full, 0, err, 235);
        const uint64_t db0 = tc::make_smem_desc_sw128(tc::smem_u32(sW), 1024);
        int i = 0, ti = 0;
        for (int tp = tp0; ok && tp < tile_pairs; tp += tpstride, ++ti) {
            const int as = ti & 1;
            if (ti >= 2 && !tc::mbar_wait_cluster(&t_empty[as], ((ti >> 1) & 1) ^ 1, err, 236)) break;
            const uint32_t d_tmem = tmem_base + (uint32_t)(as * ACC_COLS);
#pragma unroll
            for (int c = 0; c < KCH; ++c, ++i) {
                const int s = i % STAGES;
                if (!tc::mbar_wait(&a_full[s], (i / STAGES) & 1, err, 237)) { ok = false; break; }
                if (!tc::mbar_wait_cluster(&pa_full[s], (i / STAGES) & 1, err, 238)) { ok = false; break; }
                tc::tc_fence_after();
                const uint64_t da0 = tc::make_smem_desc_sw128(tc::smem_u32(sA + s * H2_HALO_STRIDE), H2_HW * 128);
                if (tc::elect_one()) {
                    constexpr int KH = KCH / 2;
                    const bool lo_chunk = X3 && c >= KH;
                    const int wpos = X3 ? 2 * (lo_chunk ? c - KH : c) : c;            // weight block: its hi rows (X3: followed by its lo rows)
                    const uint32_t id = (X3 && !lo_chunk) ? idesc2 : idesc;
                    const uint32_t d_out = d_tmem + (lo_chunk ? (uint32_t)(2 * HP_N) : 0u);
                    const int first = lo_chunk ? c - KH : c;                           // first chunk writing this accumulator region
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
#pragma unroll
                        for (int k2 = 0; k2 < 4; ++k2) {
                            const uint64_t da = da0 + (uint64_t)((((t / 3) * H2_HW + (t % 3)) * 128 + k2 * 32) >> 4);
                            const uint64_t db = db0 + (uint64_t)(((t * KCH + wpos) * W_TAP_BYTES + k2 * 32) >> 4);
                            tc::umma_bf16_pair(d_out, da, db, id, (first | t | k2) ? 1u : 0u);
                        }
                    }
                    tc::umma_commit_pair(&a_empty[s]);
                    if (c == KCH - 1) tc::umma_commit_pair(&t_full[as]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int yl = 4 * q + (lane >> 3), xl = lane & 7;
        const uint32_t leader_te[2] = {tc::map_to_cta(&t_empty[0], 0), tc::map_to_cta(&t_empty[1], 0)};
        uint8_t* stg = reinterpret_cast<uint8_t*>(s_bias + HP_N) + (warp - 4) * 2048;
        const int opitch = X3 ? 2 * cout : cout;
        int i = 0;
        for (int tp = tp0; tp < tile_pairs; tp += tpstride, ++i) {
            const int as = i & 1;
            const int tile = 2 * tp + (int)rank;
            const int img = tile / tiles_per_img, rem = tile % tiles_per_img;
            const int y = (rem / tiles_x) * H2_TH + yl, x = (rem % tiles_x) * H2_TW + xl;
            if (!tc::mbar_wait_cluster(&t_full[as], (i >> 1) & 1, err, 239)) break;
            tc::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * ACC_COLS);
            const bool inside = (tile < total) && (y < h) && (x < w);
            size_t pix;
            bool writer;
            if (pool) {
                writer = inside && ((xl | yl) & 1) == 0;
                pix = ((size_t)img * (h / 2) + (y >> 1)) * (size_t)(w / 2) + (x >> 1);
            } else {
                writer = inside;
                pix = ((size_t)img * h + y) * (size_t)w + x;
            }
#pragma unroll 1
            for (int c0 = 0; c0 < HP_N; c0 += 32) {
                uint32_t v[32];
                if constexpr (X3) {
                    // channels [c0, c0 + 32) belong to CTA (c0 / 32)'s weight rows: hh at 64 r, hl at 64 r + 32, lh at 2 NP + c0
                    uint32_t v2[32], v3[32];
                    tc::tmem_ld32(taddr + 2 * c0, v);
                    tc::tmem_ld32(taddr + 2 * c0 + 32, v2);
                    tc::tmem_ld32(taddr + 2 * HP_N + c0, v3);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        v[j] = __float_as_uint(__fadd_rn(__fadd_rn(__uint_as_float(v2[j]), __uint_as_float(v3[j])), __uint_as_float(v[j])));
                } else {
                    tc::tmem_ld32(taddr + c0, v);
                    tc::tmem_ld_wait();
                }
                if constexpr (X3) {
                    uint32_t phi[16], plo[16];
                    tc::epilogue_split32(v, &s_bias[c0], relu, pool, phi, plo);
                    if (writer) {
                        uint4* oh = reinterpret_cast<uint4*>(out_bf + pix * opitch + ch0 + c0);
                        uint4* ol = reinterpret_cast<uint4*>(out_bf + pix * opitch + cout + ch0 + c0);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            oh[g] = make_uint4(phi[4 * g], phi[4 * g + 1], phi[4 * g + 2], phi[4 * g + 3]);
                            ol[g] = make_uint4(plo[4 * g], plo[4 * g + 1], plo[4 * g + 2], plo[4 * g + 3]);
                        }
                    }
                } else {
                uint32_t packed[16];
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 bb = *reinterpret_cast<const float4*>(&s_bias[c0 + 4 * j4]);
                    const float a0 = __uint_as_float(v[4 * j4]) + bb.x, a1 = __uint_as_float(v[4 * j4 + 1]) + bb.y;
                    const float a2 = __uint_as_float(v[4 * j4 + 2]) + bb.z, a3 = __uint_as_float(v[4 * j4 + 3]) + bb.w;
                    packed[2 * j4] = relu ? tc::pack_bf16x2_relu(a0, a1) : tc::pack_bf16x2(a0, a1);
                    packed[2 * j4 + 1] = relu ? tc::pack_bf16x2_relu(a2, a3) : tc::pack_bf16x2(a2, a3);
                }
                if (pool) {  // max of bf16-rounded values == rounding of the max (monotone)
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        uint32_t u = packed[j];
                        uint32_t u1 = __shfl_xor_sync(0xffffffffu, u, 1);
                        __nv_bfloat162 o = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&u), *reinterpret_cast<__nv_bfloat162*>(&u1));
                        u = *reinterpret_cast<uint32_t*>(&o);
                        uint32_t u8 = __shfl_xor_sync(0xffffffffu, u, 8);
                        o = __hmax2(o, *reinterpret_cast<__nv_bfloat162*>(&u8));
                        packed[j] = *reinterpret_cast<uint32_t*>(&o);
                    }
                }
                if (pool) {
                    if (writer) {
                        uint4* o = reinterpret_cast<uint4*>(out_bf + pix * cout + ch0 + c0);
#pragma unroll
                        for (int g = 0; g < 4; ++g) o[g] = make_uint4(packed[4 * g], packed[4 * g + 1], packed[4 * g + 2], packed[4 * g + 3]);
                    }
                } else {
                    // full-resolution write-out through this warp's 2 KB staging tile ([32 px][64 B], pieces XOR-swizzled):
                    // four lanes then store one pixel's 64 contiguous bytes, so an instruction touches 8 lines instead of 32
                    // (global stores share the L1 / shared-memory data path the MMA operands are fetched through)
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        *reinterpret_cast<uint4*>(stg + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) =
                            make_uint4(packed[4 * g], packed[4 * g + 1], packed[4 * g + 2], packed[4 * g + 3]);
                    __syncwarp();
#pragma unroll
                    for (int it4 = 0; it4 < 4; ++it4) {
                        const int qq = lane + 32 * it4, pp = qq >> 2, j = qq & 3;
                        const int py = y - yl + 4 * q + (pp >> 3), px = x - xl + (pp & 7);
                        if (tile < total && py < h && px < w)
                            *reinterpret_cast<uint4*>(out_bf + (((size_t)img * h + py) * (size_t)w + px) * cout + ch0 + c0 + j * 8) =
                                *reinterpret_cast<const uint4*>(stg + pp * 64 + ((j ^ ((pp >> 1) & 3)) << 4));
                    }
                    __syncwarp();
                }
                }   // !X3
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_cluster(leader_te[as]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::cluster_sync_all();     // neither CTA frees TMEM / exits while the other may still signal it
    if (warp == 2) {
        tc::tc_fence_after();
        tc::tmem_dealloc_pair(tmem_base, TMEM_COLS);
    }
}

template <int KCH, int STAGES, int NP = 128, bool X3 = false>
static int launch_conv_tc_halo_pair(gnb_ctx* ctx, const CUtensorMap& tin, const CUtensorMap& tw, const ConvLayer& L, int n, int h, int w,
                                    bf16* out_bf, int relu, int pool, const char* name) {
    constexpr int smem = 1024 + 9 * KCH * (NP / 2) * 128 + STAGES * H2_HALO_STRIDE + 256 + NP * 4 + (X3 ? 0 : 4 * 2048);   // + write-out staging
    GNB_CUDA(ctx, gnb_func_smem(ctx, conv_tc_halo_pair_kernel<KCH, STAGES, NP, X3>, smem));
    const int n_slices = L.cout_pad / NP;
    const int tile_pairs = (ceil_div(w, H2_TW) * ceil_div(h, H2_TH) * n + 1) / 2;
    int per_slice = (ctx->sm_count / 2) / n_slices;
    if (per_slice > tile_pairs) per_slice = tile_pairs;
    if (per_slice < 1) per_slice = 1;
    const int grid = 2 * per_slice * n_slices;
    GNB_KERNEL(ctx, name, conv_tc_halo_pair_kernel<KCH, STAGES, NP, X3><<<grid, 256, smem, ctx->stream>>>(
        tin, tw, L.bias, h, w, n, L.cout, n_slices, out_bf, relu, pool, gnb_tc_err_dev(ctx)));
    return GNB_OK;
}


// ------------------------------------------------------------------------------------------------
// conv1a + conv1b + 2x2 max-pool in ONE kernel.  The 64-channel full-resolution activation
// (128 B/pixel written and read back = 2/3 of the whole stack's HBM traffic) never leaves the SM:
//   im2col warps : 3x3 neighbourhood of the u8 image -> bf16 [256 halo rows x K=16] operand (9 taps)
//   MMA1         : conv1a as two tcgen05.mma (M=128, N=64, K=16) -> TMEM T1 (fp32)
//   convert warps: T1 + bias, ReLU, bf16 -> the 18x10x64ch halo tile of conv1b, written directly in
//                  the 128B-swizzle layout (zero outside the image = conv1b's padding)
//   MMA2         : conv1b, 36 tcgen05.mma over nine shifted descriptor views (as conv_tc_halo_kernel)
//   epilogue     : bias, ReLU, 2x2 max-pool, bf16 store at half resolution
// All five stages are double-buffered and run concurrently on different tiles.
#define F1_A1_BYTES (256 * 128)      // 256 rows x 128 B (only the first 32 B = K 16 of a row are used)
#define F1_W1A_BYTES (64 * 128)
#define F1_W1B_BYTES (9 * 64 * 128)
#define F1_THREADS 512               // 16 warps: 0 W-loader, 1 MMA, 2 TMEM alloc, 4-7 epilogue, 8-11 convert, 12-15 im2col
#define F1_PATCH_W 32                 // bytes per patch row (TMA inner box extent must be a multiple of 16 B)
#define F1_PATCH_H (H2_HH + 2)        // 20 rows: halo of the halo
#define F1_PATCH_BYTES (F1_PATCH_W * F1_PATCH_H)   // 640
#define F1_PSTAGES 8
#define F1_SMEM (1024 + F1_W1B_BYTES + F1_W1A_BYTES + 2 * F1_A1_BYTES + 2 * H2_HALO_STRIDE + 1024 + F1_PSTAGES * F1_PATCH_BYTES)

__global__ void __launch_bounds__(F1_THREADS, 1) conv1_fused_kernel(const __grid_constant__ CUtensorMap tmap_img, const bf16* __restrict__ w1a,
                                                                    const float* __restrict__ bias1a,
                                                                    const __grid_constant__ CUtensorMap tmap_w1b,
                                                                    const float* __restrict__ bias1b, int h, int w, int n_img,
                                                                    bf16* __restrict__ out_bf, int* err) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW1b = smem;
    uint8_t* sW1a = sW1b + F1_W1B_BYTES;
    uint8_t* sA1 = sW1a + F1_W1A_BYTES;                 // [2][256 x 128 B]
    uint8_t* sA2 = sA1 + 2 * F1_A1_BYTES;               // [2][halo]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA2 + 2 * H2_HALO_STRIDE);
    uint64_t* w_full = bars;          // 1
    uint64_t* a1_full = bars + 1;     // [2] im2col -> MMA1      (4 arrivals: one per im2col warp)
    uint64_t* a1_empty = bars + 3;    // [2] MMA1 done reading A1 (tcgen05.commit)
    uint64_t* t1_full = bars + 5;     // [2] MMA1 -> convert      (tcgen05.commit)
    uint64_t* t1_empty = bars + 7;    // [2] convert -> MMA1      (4 arrivals)
    uint64_t* a2_full = bars + 9;     // [2] convert -> MMA2      (4 arrivals)
    uint64_t* a2_empty = bars + 11;   // [2] MMA2 done reading A2 (tcgen05.commit)
    uint64_t* t2_full = bars + 13;    // [2] MMA2 -> epilogue     (tcgen05.commit)
    uint64_t* t2_empty = bars + 15;   // [2] epilogue -> MMA2     (4 arrivals)
    uint64_t* p_full = bars + 17;     // [F1_PSTAGES] TMA image patch landed
    uint64_t* p_empty = bars + 17 + F1_PSTAGES;  // [F1_PSTAGES] im2col done with the patch (4 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17 + 2 * F1_PSTAGES);
    float* s_b1a = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);
    float* s_b1b = s_b1a + 64;
    uint8_t* sP = reinterpret_cast<uint8_t*>(bars) + 1024;   // [F1_PSTAGES][20 x 32 B] u8 image patches
    __shared__ unsigned short s_lut[256];   // u8 -> bf16(v / 255)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (w + H2_TW - 1) / H2_TW, tiles_y = (h + H2_TH - 1) / H2_TH;
    const int tiles_per_img = tiles_x * tiles_y;
    const int total = tiles_per_img * n_img;
    if (threadIdx.x >= 256) {
        const __nv_bfloat16 q = __float2bfloat16_rn((float)(threadIdx.x - 256) / 255.0f);
        s_lut[threadIdx.x - 256] = *reinterpret_cast<const unsigned short*>(&q);
    }

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmap_w1b);
        tc::tma_prefetch_desc(&tmap_img);
        tc::mbar_init(w_full, 1);
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&a1_full[s], 4); tc::mbar_init(&a1_empty[s], 1);
            tc::mbar_init(&t1_full[s], 1); tc::mbar_init(&t1_empty[s], 4);
            tc::mbar_init(&a2_full[s], 4); tc::mbar_init(&a2_empty[s], 1);
            tc::mbar_init(&t2_full[s], 1); tc::mbar_init(&t2_empty[s], 4);
        }
        for (int s = 0; s < F1_PSTAGES; ++s) { tc::mbar_init(&p_full[s], 1); tc::mbar_init(&p_empty[s], 4); }
        tc::fence_barrier_init();
    }
    if (warp == 2) { tc::tmem_alloc(tmem_slot, 512); tc::tmem_relinquish(); }
    if (threadIdx.x < 64) { s_b1a[threadIdx.x] = bias1a[threadIdx.x]; s_b1b[threadIdx.x] = bias1b[threadIdx.x]; }
    // conv1a weights as the B operand of MMA1: row n (output channel), K = 16 (taps 0..8, then zeros)
    if (threadIdx.x >= 64 && threadIdx.x < 128) {
        const int nrow = threadIdx.x - 64;
        __align__(16) bf16 k[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) k[t] = t < 9 ? w1a[t * 64 + nrow] : __float2bfloat16_rn(0.f);
        // K slots 9 and 10 multiply a constant 1.0 in the A operand: bias = hi + lo (two bf16 terms,
        // exact to 2^-17 relative), accumulated in fp32 by the tensor core
        const float bfull = bias1a[nrow];
        k[9] = __float2bfloat16_rn(bfull);
        k[10] = __float2bfloat16_rn(bfull - __bfloat162float(k[9]));
        uint8_t* row = sW1a + nrow * 128;
        *reinterpret_cast<uint4*>(row + ((0 ^ (nrow & 7)) << 4)) = *reinterpret_cast<const uint4*>(&k[0]);
        *reinterpret_cast<uint4*>(row + ((1 ^ (nrow & 7)) << 4)) = *reinterpret_cast<const uint4*>(&k[8]);
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns: T1 buffers at 0 and 128 (two 64-col halves each: rows 0-127 / 128-255), T2 at 256 and 320

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(w_full, F1_W1B_BYTES);
            for (int t = 0; t < 9; ++t) tc::tma_load_3d(sW1b + t * 64 * 128, &tmap_w1b, w_full, 0, 0, t);
            // u8 image patches (20 rows x 32 px, origin (y0-2, x0-2); OOB zero fill = image zero padding)
            int i = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
                const int s = i % F1_PSTAGES;
                if (i >= F1_PSTAGES && !tc::mbar_wait(&p_empty[s], ((i / F1_PSTAGES) & 1) ^ 1, err, 410)) break;
                const int im = tile / tiles_per_img, rem = tile % tiles_per_img;
                const int y0 = (rem / tiles_x) * H2_TH, x0 = (rem % tiles_x) * H2_TW;
                tc::mbar_arrive_expect_tx(&p_full[s], F1_PATCH_BYTES);
                // the image is viewed as uint32 words; the box starts on a 16-byte boundary at or left of x0-2
                const int sx = ((x0 - 2 + 16) & ~15) - 16;
                tc::tma_load_3d(sP + s * F1_PATCH_BYTES, &tmap_img, &p_full[s], sx >> 2, y0 - 2, im);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: iteration i issues MMA1(i) then MMA2(i-1) =====
        const uint32_t idesc = tc::make_idesc_bf16(128, 64);
        bool ok = tc::mbar_wait(w_full, 0, err, 401);
        const uint64_t dw1a = tc::make_smem_desc_sw128(tc::smem_u32(sW1a), 1024);
        const uint64_t dw1b = tc::make_smem_desc_sw128(tc::smem_u32(sW1b), 1024);
        const uint64_t da1_0 = tc::make_smem_desc_sw128(tc::smem_u32(sA1), 1024);
        const uint64_t da2_0 = tc::make_smem_desc_sw128(tc::smem_u32(sA2), H2_HW * 128);
        const int my_tiles = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        for (int i = 0; ok && i <= my_tiles; ++i) {
            if (i < my_tiles) {
                const int b = i & 1;
                const uint32_t ph = (i >> 1) & 1;
                if (!tc::mbar_wait(&a1_full[b], ph, err, 402)) break;
                if (i >= 2 && !tc::mbar_wait(&t1_empty[b], ph ^ 1, err, 403)) break;
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint64_t da = da1_0 + (uint64_t)((b * F1_A1_BYTES) >> 4);
                    tc::umma_bf16(tmem_base + (uint32_t)(b * 128), da, dw1a, idesc, 0u);
                    tc::umma_bf16(tmem_base + (uint32_t)(b * 128 + 64), da + (uint64_t)((128 * 128) >> 4), dw1a, idesc, 0u);
                    tc::umma_commit(&a1_empty[b]);
                    tc::umma_commit(&t1_full[b]);
                }
                __syncwarp();
            }
            if (i >= 1) {
                const int j = i - 1, b = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                if (!tc::mbar_wait(&a2_full[b], ph, err, 404)) break;
                if (j >= 2 && !tc::mbar_wait(&t2_empty[b], ph ^ 1, err, 405)) break;
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint64_t da0 = da2_0 + (uint64_t)((b * H2_HALO_STRIDE) >> 4);
                    const uint32_t d_tmem = tmem_base + 256u + (uint32_t)(b * 64);
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t da = da0 + (uint64_t)((((t / 3) * H2_HW + (t % 3)) * 128 + k * 32) >> 4);
                            const uint64_t db = dw1b + (uint64_t)((t * 64 * 128 + k * 32) >> 4);
                            tc::umma_bf16(d_tmem, da, db, idesc, (t | k) ? 1u : 0u);
                        }
                    }
                    tc::umma_commit(&a2_empty[b]);
                    tc::umma_commit(&t2_full[b]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 12) {
        // ===== im2col: halo row r = hy*10 + hx  <->  a1a pixel (y0-1+hy, x0-1+hx); K index = tap =====
        const int it_ = threadIdx.x - 12 * 32;  // 0..127
        // rows owned by this thread (tile independent): r0 = it_, r1 = it_ + 128 (< 180 for it_ < 52)
        const int r0 = it_, r1 = it_ + 128;
        const int hy0 = r0 / H2_HW, hx0 = r0 - hy0 * H2_HW, hy1 = r1 / H2_HW, hx1 = r1 - hy1 * H2_HW;
        const bool has1 = r1 < H2_HH * H2_HW;
        const unsigned short one = 0x3F80;   // bf16 1.0
        int i = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
            const int b = i & 1, ps = i % F1_PSTAGES;
            if (i >= 2 && !tc::mbar_wait(&a1_empty[b], ((i >> 1) & 1) ^ 1, err, 406)) break;
            if (!tc::mbar_wait(&p_full[ps], (i / F1_PSTAGES) & 1, err, 411)) break;
            const int rem = tile % tiles_per_img;
            const int y0 = (rem / tiles_x) * H2_TH, x0 = (rem % tiles_x) * H2_TW;
            const uint8_t* patch = sP + ps * F1_PATCH_BYTES + ((x0 - 2) - (((x0 - 2 + 16) & ~15) - 16));  // column of x0-2 in the box
            uint8_t* a1 = sA1 + b * F1_A1_BYTES;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                if (rr == 1 && !has1) break;
                const int r = rr ? r1 : r0;
                const int hy = rr ? hy1 : hy0, hx = rr ? hx1 : hx0;
                const int y = y0 - 1 + hy, x = x0 - 1 + hx;
                uint32_t kw[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // 16 bf16 = K slots 0..15
                if (y >= 0 && y < h && x >= 0 && x < w) {   // outside the image the whole row stays 0 => conv1a output 0
                    unsigned short kv[9];
#pragma unroll
                    for (int t = 0; t < 9; ++t) kv[t] = s_lut[patch[(hy + t / 3) * F1_PATCH_W + hx + t % 3]];
                    kw[0] = kv[0] | ((uint32_t)kv[1] << 16); kw[1] = kv[2] | ((uint32_t)kv[3] << 16);
                    kw[2] = kv[4] | ((uint32_t)kv[5] << 16); kw[3] = kv[6] | ((uint32_t)kv[7] << 16);
                    kw[4] = kv[8] | ((uint32_t)one << 16);   // slot 9 = 1.0 (bias hi)
                    kw[5] = one;                             // slot 10 = 1.0 (bias lo)
                }
                uint8_t* row = a1 + r * 128;
                *reinterpret_cast<uint4*>(row + ((0 ^ (r & 7)) << 4)) = make_uint4(kw[0], kw[1], kw[2], kw[3]);
                *reinterpret_cast<uint4*>(row + ((1 ^ (r & 7)) << 4)) = make_uint4(kw[4], kw[5], kw[6], kw[7]);
            }
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) { tc::mbar_arrive(&a1_full[b]); tc::mbar_arrive(&p_empty[ps]); }
        }
    } else if (warp >= 8) {
        // ===== convert: T1 (fp32 conv1a) -> bias, ReLU, bf16 -> conv1b halo tile (128B swizzle) =====
        const int q = warp & 3;
        int i = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
            const int b = i & 1;
            const uint32_t ph = (i >> 1) & 1;
            if (!tc::mbar_wait(&t1_full[b], ph, err, 407)) break;
            if (i >= 2 && !tc::mbar_wait(&a2_empty[b], ph ^ 1, err, 408)) break;
            tc::tc_fence_after();
            uint8_t* a2 = sA2 + b * H2_HALO_STRIDE;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                if (half * 128 + q * 32 >= H2_HH * H2_HW) continue;  // rows 192..255 hold no halo pixel (warp-uniform)
                const int r = half * 128 + q * 32 + lane;       // halo row handled by this thread
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 128 + half * 64);
                uint8_t* row = a2 + r * 128;
#pragma unroll 1
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    uint32_t v[32];
                    tc::tmem_ld32(taddr + c0, v);
                    tc::tmem_ld_wait();
                    if (r < H2_HH * H2_HW) {   // bias came through the MMA; out-of-image rows are exactly 0
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int chunk = (c0 >> 3) + g;   // 16-byte chunk index 0..7
                            *reinterpret_cast<uint4*>(row + ((chunk ^ (r & 7)) << 4)) = make_uint4(
                                tc::pack_bf16x2_relu(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1])),
                                tc::pack_bf16x2_relu(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])),
                                tc::pack_bf16x2_relu(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])),
                                tc::pack_bf16x2_relu(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])));
                        }
                    }
                }
            }
            tc::tc_fence_before();
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) { tc::mbar_arrive(&t1_empty[b]); tc::mbar_arrive(&a2_full[b]); }
        }
    } else if (warp >= 4) {
        // ===== epilogue (as conv_tc_halo_kernel, pool fixed on) =====
        const int q = warp & 3;
        const int yl = 4 * q + (lane >> 3), xl = lane & 7;
        int i = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
            const int b = i & 1;
            const int im = tile / tiles_per_img, rem = tile % tiles_per_img;
            const int y = (rem / tiles_x) * H2_TH + yl, x = (rem % tiles_x) * H2_TW + xl;
            if (!tc::mbar_wait(&t2_full[b], (i >> 1) & 1, err, 409)) break;
            tc::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 256u + (uint32_t)(b * 64);
            const bool writer = (y < h) && (x < w) && ((xl | yl) & 1) == 0;
            const size_t pix = ((size_t)im * (h / 2) + (y >> 1)) * (size_t)(w / 2) + (x >> 1);
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t v[32];
                tc::tmem_ld32(taddr + c0, v);
                tc::tmem_ld_wait();
                uint32_t packed[16];
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 bb = *reinterpret_cast<const float4*>(&s_b1b[c0 + 4 * j4]);
                    packed[2 * j4] = tc::pack_bf16x2_relu(__uint_as_float(v[4 * j4]) + bb.x, __uint_as_float(v[4 * j4 + 1]) + bb.y);
                    packed[2 * j4 + 1] = tc::pack_bf16x2_relu(__uint_as_float(v[4 * j4 + 2]) + bb.z, __uint_as_float(v[4 * j4 + 3]) + bb.w);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    uint32_t u = packed[j];
                    uint32_t u1 = __shfl_xor_sync(0xffffffffu, u, 1);
                    __nv_bfloat162 o = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&u), *reinterpret_cast<__nv_bfloat162*>(&u1));
                    u = *reinterpret_cast<uint32_t*>(&o);
                    uint32_t u8 = __shfl_xor_sync(0xffffffffu, u, 8);
                    o = __hmax2(o, *reinterpret_cast<__nv_bfloat162*>(&u8));
                    packed[j] = *reinterpret_cast<uint32_t*>(&o);
                }
                if (writer) {
                    uint4* o = reinterpret_cast<uint4*>(out_bf + pix * 64 + c0);
#pragma unroll
                    for (int g = 0; g < 4; ++g) o[g] = make_uint4(packed[4 * g], packed[4 * g + 1], packed[4 * g + 2], packed[4 * g + 3]);
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&t2_empty[b]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}

int gnb_conv1_fused_tc(gnb_ctx* ctx, const uint8_t* img, int n, int h, int w, bf16* out_p1) {
    if (((h | w) & 1) || (w % 16)) return GNB_E_INVALID;   // TMA: row pitch must be a multiple of 16 bytes
    GNB_CUDA(ctx, gnb_func_smem(ctx, conv1_fused_kernel, F1_SMEM));
    gnb_encode_tiled_fn fn = gnb_get_encode_tiled(ctx);
    if (!fn) return GNB_E_CUDA;
    CUtensorMap timg;
    const cuuint64_t gdim[3] = {(cuuint64_t)w / 4, (cuuint64_t)h, (cuuint64_t)n};
    const cuuint64_t gstr[2] = {(cuuint64_t)w, (cuuint64_t)w * h};
    const cuuint32_t box[3] = {F1_PATCH_W / 4, F1_PATCH_H, 1}, es[3] = {1, 1, 1};
    CUresult r = fn(&timg, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(img), gdim, gstr, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { GNB_SET_ERR(ctx, "cuTensorMapEncodeTiled(u8 image) failed: %d", (int)r); return GNB_E_CUDA; }
    const int total = ceil_div(w, H2_TW) * ceil_div(h, H2_TH) * n;
    const int grid = total < ctx->sm_count ? total : ctx->sm_count;
    GNB_KERNEL(ctx, "conv_tc:1a+1b", conv1_fused_kernel<<<grid, F1_THREADS, F1_SMEM, ctx->stream>>>(
        timg, ctx->layers[L1A].w, ctx->layers[L1A].bias, tc_state(ctx)->layers[L1B].w64, ctx->layers[L1B].bias, h, w, n, out_p1, gnb_tc_err_dev(ctx)));
    return GNB_OK;
}

template <int NPAD>
static int launch_conv_tc(gnb_ctx* ctx, const CUtensorMap& tin, const CUtensorMap& tw, const ConvLayer& L, int n, int h, int w,
                          bf16* out_bf, float* out_f, int relu, int pool, const char* name) {
    constexpr int smem = CT_STAGES * (CT_A_BYTES + NPAD * 128) + 1024 + 256;
    GNB_CUDA(ctx, gnb_func_smem(ctx, conv_tc_kernel<NPAD>, smem));
    dim3 grid(ceil_div(w, CT_TW), ceil_div(h, CT_TH), n);
    GNB_KERNEL(ctx, name, conv_tc_kernel<NPAD><<<grid, 256, smem, ctx->stream>>>(
        tin, tw, L.bias, h, w, L.cin, L.ks, L.cout, out_bf, out_f, relu, pool, gnb_tc_err_dev(ctx)));
    return GNB_OK;
}

const CUtensorMap* gnb_conv_tc_wmap(gnb_ctx* ctx, int lid) {
    TcLayerMaps& m = tc_state(ctx)->layers[lid];
    return m.valid ? &m.w : nullptr;
}
const CUtensorMap* gnb_conv_tc_wmap_x3(gnb_ctx* ctx, int lid) {   // split (hi | lo) weights of a 1x1 head
    TcLayerMaps& m = tc_state(ctx)->layers[lid];
    return m.valid ? &m.x64 : nullptr;
}

int gnb_conv_tc_init(gnb_ctx* ctx) {
    if (!gnb_tc_err_dev(ctx)) { GNB_SET_ERR(ctx, "cannot allocate the host-mapped error word"); return GNB_E_CUDA; }
    TcLayerMaps* g_wmaps = tc_state(ctx)->layers;
    for (int l = 0; l < GNB_NUM_LAYERS; ++l) {
        const ConvLayer& L = ctx->layers[l];
        g_wmaps[l].valid = 0;
        if (L.cin % 64) continue;  // conv1a (Cin = 1) stays on the CUDA cores
        const uint64_t dims[3] = {(uint64_t)L.cin, (uint64_t)L.cout_pad, (uint64_t)(L.ks * L.ks)};
        const uint64_t strides[2] = {(uint64_t)L.cin * 2, (uint64_t)L.cout_pad * L.cin * 2};
        const uint32_t box[3] = {64, (uint32_t)L.cout_pad, 1};
        int rc = gnb_make_tmap_bf16(ctx, &g_wmaps[l].w, L.w, 3, dims, strides, box);
        if (rc) return rc;
        const uint32_t box64[3] = {64, 64, 1};
        if ((rc = gnb_make_tmap_bf16(ctx, &g_wmaps[l].w64, L.w, 3, dims, strides, box64))) return rc;
        if (L.cout_pad >= 128) {
            const uint32_t box128[3] = {64, 128, 1};
            if ((rc = gnb_make_tmap_bf16(ctx, &g_wmaps[l].w128, L.w, 3, dims, strides, box128))) return rc;
        }
        if (ctx->cfg.precision == 1 && L.ks == 1) {
            // 1x1 heads: split weights [1][cout_pad][hi: cin | lo: cin], all output rows in one box
            const uint64_t xd[3] = {(uint64_t)L.cin * 2, (uint64_t)L.cout_pad, 1};
            const uint64_t xs[2] = {(uint64_t)L.cin * 4, (uint64_t)L.cout_pad * L.cin * 4};
            const uint32_t xb[3] = {64, (uint32_t)L.cout_pad, 1};
            if ((rc = gnb_make_tmap_bf16(ctx, &g_wmaps[l].x64, L.w_x3, 3, xd, xs, xb))) return rc;
        }
        if (ctx->cfg.precision == 1 && L.ks == 3) {
            // split weights [tap][cout_pad][hi: cin | lo: cin]: chunk c of the inner dimension is a hi chunk for
            // c < cin / 64 and the matching lo chunk after that
            const uint64_t xd[3] = {(uint64_t)L.cin * 2, (uint64_t)L.cout_pad, (uint64_t)(L.ks * L.ks)};
            const uint64_t xs[2] = {(uint64_t)L.cin * 4, (uint64_t)L.cout_pad * L.cin * 4};
            const uint32_t xb64[3] = {64, 64, 1}, xb32[3] = {64, 32, 1};
            if ((rc = gnb_make_tmap_bf16(ctx, &g_wmaps[l].x64, L.w_x3, 3, xd, xs, xb64))) return rc;
            if ((rc = gnb_make_tmap_bf16(ctx, &g_wmaps[l].x32, L.w_x3, 3, xd, xs, xb32))) return rc;
        }
        g_wmaps[l].valid = 1;
    }
    return GNB_OK;
}

int gnb_conv_tc_layer(gnb_ctx* ctx, const ConvLayer& L, const bf16* in, int n, int h, int w, bf16* out_bf, float* out_f,
                      int relu, int pool) {
    const int lid = (int)(&L - ctx->layers);
    TcLayerMaps* g_wmaps = tc_state(ctx)->layers;
    if (lid < 0 || lid >= GNB_NUM_LAYERS || !g_wmaps[lid].valid) return GNB_E_INVALID;
    if (pool && ((h | w) & 1)) return GNB_E_INVALID;
    static const char* kNames[GNB_NUM_LAYERS] = {"conv1a", "conv_tc:1b", "conv_tc:2a", "conv_tc:2b", "conv_tc:3a", "conv_tc:3b",
                                                 "conv_tc:4a", "conv_tc:4b", "conv_tc:Pa", "conv_tc:Pb", "conv_tc:Da", "conv_tc:Db"};
    static const int v1_only = getenv("GNB_CONV_TC_V1") ? atoi(getenv("GNB_CONV_TC_V1")) : 0;
    CUtensorMap tin;
    int rc;
    if (ctx->cfg.precision == 1) {
        // fp32-faithful mode: split activations [n][h][w][hi: cin | lo: cin], three MMAs per product (X3 kernels)
        if (L.ks != 3 || !out_bf || (L.cin != 64 && L.cin != 128) || (L.cout_pad % 64)) return GNB_E_INVALID;
        const uint64_t xd[4] = {(uint64_t)L.cin * 2, (uint64_t)w, (uint64_t)h, (uint64_t)n};
        const uint64_t xs[3] = {(uint64_t)L.cin * 4, (uint64_t)w * L.cin * 4, (uint64_t)h * w * L.cin * 4};
        const uint32_t hbox[4] = {64, H2_HW, H2_HH, 1};
        if ((rc = gnb_make_tmap_bf16(ctx, &tin, const_cast<bf16*>(in), 4, xd, xs, hbox))) return rc;
        static const char* kNamesX3[GNB_NUM_LAYERS] = {"conv1a", "conv_x3:1b", "conv_x3:2a", "conv_x3:2b", "conv_x3:3a", "conv_x3:3b",
                                                       "conv_x3:4a", "conv_x3:4b", "conv_x3:Pa", "conv_x3:Pb", "conv_x3:Da", "conv_x3:Db"};
        if (L.cin == 64)    // single CTAs, 64-channel output slices, hi + lo weights of a slice resident (144 KB)
            return launch_conv_tc_halo<64, 2, 3, false, true>(ctx, tin, g_wmaps[lid].x64, L, n, h, w, out_bf, relu, pool, kNamesX3[lid]);
        // Cin = 128: CTA pairs, M = 256 x N = 64, each CTA keeps 32 rows of the hi + lo weights of all four chunks (144 KB)
        return launch_conv_tc_halo_pair<4, 3, 64, true>(ctx, tin, g_wmaps[lid].x32, L, n, h, w, out_bf, relu, pool, kNamesX3[lid]);
    }
    const uint64_t dims[4] = {(uint64_t)L.cin, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)L.cin * 2, (uint64_t)w * L.cin * 2, (uint64_t)h * w * L.cin * 2};
    if (!v1_only && L.ks == 3 && out_bf && (L.cin == 64 || L.cin == 128) && (L.cout_pad % 64) == 0) {
        const uint32_t hbox[4] = {64, H2_HW, H2_HH, 1};
        if ((rc = gnb_make_tmap_bf16(ctx, &tin, const_cast<bf16*>(in), 4, dims, strides, hbox))) return rc;
        if (L.cin == 64 && L.cout_pad == 64)
            return launch_conv_tc_halo<64, 1, 4>(ctx, tin, g_wmaps[lid].w64, L, n, h, w, out_bf, relu, pool, kNames[lid]);
        if (L.cin == 64 && L.cout_pad == 128)
            return launch_conv_tc_halo<128, 1, 2, true>(ctx, tin, g_wmaps[lid].w128, L, n, h, w, out_bf, relu, pool, kNames[lid]);
        static const int no_pair = getenv("GNB_CONV_NO_PAIR") ? atoi(getenv("GNB_CONV_NO_PAIR")) : 0;
        if (L.cin == 128 && !no_pair && (L.cout_pad % 128) == 0)   // CTA pairs: M = 256, N = 128 MMAs, half of the weight rows per SM
            return launch_conv_tc_halo_pair<2, 3>(ctx, tin, g_wmaps[lid].w64, L, n, h, w, out_bf, relu, pool, kNames[lid]);
        if (L.cin == 128)  // 128 -> 128 / 256: slices of 64 output channels, weights of a slice resident (144 KB)
            return launch_conv_tc_halo<64, 2, 3>(ctx, tin, g_wmaps[lid].w64, L, n, h, w, out_bf, relu, pool, kNames[lid]);
    }
    const uint32_t box[4] = {64, CT_TW, CT_TH, 1};
    rc = gnb_make_tmap_bf16(ctx, &tin, const_cast<bf16*>(in), 4, dims, strides, box);
    if (rc) return rc;
    switch (L.cout_pad) {
        case 64: return launch_conv_tc<64>(ctx, tin, g_wmaps[lid].w, L, n, h, w, out_bf, out_f, relu, pool, kNames[lid]);
        case 96: return launch_conv_tc<96>(ctx, tin, g_wmaps[lid].w, L, n, h, w, out_bf, out_f, relu, pool, kNames[lid]);
        case 128: return launch_conv_tc<128>(ctx, tin, g_wmaps[lid].w, L, n, h, w, out_bf, out_f, relu, pool, kNames[lid]);
        case 256: return launch_conv_tc<256>(ctx, tin, g_wmaps[lid].w, L, n, h, w, out_bf, out_f, relu, pool, kNames[lid]);
        default: return GNB_E_INVALID;
    }
}
