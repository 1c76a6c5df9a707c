// lightglue.cu — the LightGlue transformer layers in front of the assignment head (SURVEY.md §8(f)
// rank 1): what the reference matcher actually runs before K4,
//   LightGlueMatcher("sift", {"n_layers": 9, "depth_confidence": -1, "width_confidence": -1, ...})
//   ros/gisnav/gisnav/core/pose_node.py:109-121, called at :285-287
// restated from the published model (oracle/lightglue_ref.py, pinned against transformers 5.5).
//
// Per layer and image: self block (q,k,v projections, rotary position encoding from the keypoint
// coordinates, 4-head attention, output projection, MLP [x, o] -> 512 -> LayerNorm -> GELU -> 256,
// residual add) then a cross block (same shape, keys/values from the other image, no rotation).
//
// Everything dense runs on tcgen05 with bf16 operands and fp32 TMEM accumulators:
//   lg_linear_persist_kernel<MODE>  Y = A W^T + b with the weight slice resident in shared memory and the
//                           token tiles streamed by TMA (persistent CTAs, two TMEM accumulator stages);
//                           the epilogue (one token row per thread, tcgen05.ld) is specialised per use:
//                             QKV  bias, rotary, 1/sqrt(64) folded into q, V written transposed
//                             OUT  bias -> second half of the MLP input row
//                             FC2  bias + residual add into the fp32 stream, bf16 copy for the next GEMM
//   lg_linear_kernel<FC1>   one 128-token tile per CTA, all 512 outputs in TMEM (BN = 512 = the whole TMEM
//                           width): bias, LayerNorm(512) and exact GELU straight from the resident row
//   lg_attn_kernel          two-pass softmax attention per (128 queries, head): pass 0 takes the row max of
//                           S = q k^T from TMEM; pass 1 recomputes S, writes P = exp(S - max) as the bf16 A
//                           operand (128B-swizzled smem) of the second MMA O += P V and keeps the fp32 row
//                           sum that normalises O at the end.  S and P never touch HBM.
// The fp32 residual stream is ctx->desc_f32 itself (in place), so the head (project_tc) runs unchanged.
#include "tc_common.cuh"

#include <math.h>
#include <vector>

int* gnb_tc_err_dev(gnb_ctx* ctx);

#define LG_DIM 256
#define LG_HID 512
#define LG_HEADS 4
#define LG_HD 64

enum { LG_QKV = 0, LG_OUT = 1, LG_FC1 = 2, LG_FC2 = 3 };

struct LgBlockW {
    bf16 *wqkv, *wo, *w1, *w2;             // [768][256], [256][256], [512][512], [256][512]
    float *bqkv, *bo, *b1, *lng, *lnb, *b2;
    CUtensorMap m_qkv, m_o, m_w1, m_w2;    // box {64, 256}; m_w2: box {64, 128} (FC2 runs as two 128-column slices)
};

struct LgState {
    int n_layers;
    std::vector<LgBlockW> blocks;          // [layer][self, cross]
    float pos_w[64];                        // [32][2]
    void* wblob;                            // one device allocation holding every weight
    // activations, [slots][K][...]
    bf16 *xo, *q, *k, *vt, *att, *h;
    float* cs;                              // [slots][64][K]: cos[32], sin[32] rows, token-contiguous
    CUtensorMap m_x, m_xo, m_att, m_h, m_q, m_k, m_vt;
};

static inline LgState* lg_state(gnb_ctx* ctx) { return static_cast<LgState*>(ctx->lg_state); }

struct LgPairs { int slot_a0, slot_b0, pairs; };
__device__ __forceinline__ void lg_slots(const LgPairs& p, int z, int& slot, int& partner) {
    if (z < p.pairs) { slot = p.slot_a0 + z; partner = p.slot_b0 + z; }
    else { slot = p.slot_b0 + z - p.pairs; partner = p.slot_a0 + z - p.pairs; }
}

// ---- init: bf16 copy of the descriptors + rotary tables -------------------------------------------------
struct LgPos { float w[64]; };
__global__ void __launch_bounds__(256) lg_init_kernel(const float* __restrict__ desc, const float* __restrict__ kp_xy,
                                                      const int* __restrict__ kp_count, LgPairs pr, int k_cap, LgPos pos,
                                                      float ha, float wa, float hb, float wb, bf16* __restrict__ xo,
                                                      float* __restrict__ cs) {
    int slot, partner;
    lg_slots(pr, blockIdx.y, slot, partner);
    const int n = max(kp_count[slot], 0);
    if (n == 0 || max(kp_count[partner], 0) == 0) return;
    const int row0 = blockIdx.x * 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = row0 + warp;
    if (row >= n) return;
    const size_t tok = (size_t)slot * k_cap + row;
    // 256 floats -> 256 bf16: lane handles 8 consecutive channels
    const float4 a = __ldg(reinterpret_cast<const float4*>(desc + tok * LG_DIM + lane * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(desc + tok * LG_DIM + lane * 8 + 4));
    *reinterpret_cast<uint4*>(xo + tok * LG_HID + lane * 8) =
        make_uint4(tc::pack_bf16x2(a.x, a.y), tc::pack_bf16x2(a.z, a.w), tc::pack_bf16x2(b.x, b.y), tc::pack_bf16x2(b.z, b.w));
    // rotary angle i = lane: w[i] . normalised keypoint
    const bool is_a = blockIdx.y < pr.pairs;
    const float h = is_a ? ha : hb, w = is_a ? wa : wb;
    const float scale = __fdiv_rn(fmaxf(w, h), 2.0f);
    const float x = __fdiv_rn(__fsub_rn(kp_xy[tok * 2], __fdiv_rn(w, 2.0f)), scale);
    const float y = __fdiv_rn(__fsub_rn(kp_xy[tok * 2 + 1], __fdiv_rn(h, 2.0f)), scale);
    const float ang = __fadd_rn(__fmul_rn(x, pos.w[lane * 2]), __fmul_rn(y, pos.w[lane * 2 + 1]));
    float s, c;
    sincosf(ang, &s, &c);
    // transposed table [slot][cos 0..31, sin 0..31][token]: the row-per-thread epilogue of the q/k projection then
    // reads it with consecutive lanes on consecutive tokens (one wavefront per load)
    cs[((size_t)slot * 64 + lane) * k_cap + row] = c;
    cs[((size_t)slot * 64 + 32 + lane) * k_cap + row] = s;
}

// erf to ~3e-7 absolute (Abramowitz & Stegun 7.1.26): two MUFU (rcp, ex2) + 8 FMA instead of the ~25-instruction
// branchy erff — the fc1 epilogue evaluates 16.8 M GELUs per launch and was bound by it.  The error is four orders of
// magnitude below the bf16 rounding of the result.
__device__ __forceinline__ float lg_erf_fast(float x) {
    const float ax = fabsf(x);
    const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    p *= t;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-ax * ax * 1.4426950408889634f));
    return copysignf(fmaf(-p, e, 1.0f), x);
}

// ---- linear layers ------------------------------------------------------------------------------------------
#define LGL_STAGES 2
template <int BN> struct LglCfg {
    static constexpr int A_BYTES = 128 * 128;           // 128 rows x 64 bf16
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE = A_BYTES + B_BYTES;
    static constexpr int SMEM = 1024 + LGL_STAGES * STAGE + 256 + 4 * 128 * 4 + 3 * BN * 4 + 8 * 4096;   // + LN partials + bias/gamma/beta + staging
};

// One 128-token tile per CTA, weights streamed in 64-wide K chunks.  Only fc1 uses it (MODE = LG_FC1, BN = 512: the
// LayerNorm needs all 512 outputs of a token in one CTA, and a 512 x 512 weight matrix cannot stay resident).
// threads: warp 0 TMA, 1 MMA, 2 TMEM alloc, 4-11 epilogue, group g owning columns [256 g, 256 g + 256).
#define LGL_THREADS(BN) ((BN) == 512 ? 384 : 256)
template <int MODE, int KIN, int BN>
__global__ void __launch_bounds__(LGL_THREADS(BN), (BN == 256 ? 2 : 1))
lg_linear_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const int* __restrict__ kp_count,
                 LgPairs pr, int k_cap, const float* __restrict__ bias, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                 int rotary, const float* __restrict__ cs, bf16* __restrict__ out0, bf16* __restrict__ out1, bf16* __restrict__ out2,
                 float* __restrict__ xres, int* err) {
    typedef LglCfg<BN> Cfg;
    int slot, partner;
    lg_slots(pr, blockIdx.z, slot, partner);
    const int n = max(kp_count[slot], 0);
    const int r0 = blockIdx.x * 128;
    if (r0 >= n || max(kp_count[partner], 0) == 0) return;   // uniform exit before any barrier / TMEM allocation
    const int n0 = blockIdx.y * BN;                           // first output column (= weight row) of this CTA

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LGL_STAGES * Cfg::STAGE);
    uint64_t* full = bars;                 // [2]
    uint64_t* empty = bars + 2;            // [2]
    uint64_t* acc_full = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
    float* s_stat = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // FC1: [2 stats][2 groups][128 rows]
    float* s_par = s_stat + 4 * 128;                                                     // [bias | gamma | beta][BN]
    uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_par + 3 * BN);                       // 8 x 4 KB, one per epilogue warp
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&map_a);
        tc::tma_prefetch_desc(&map_w);
        for (int s = 0; s < LGL_STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(acc_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) { tc::tmem_alloc(tmem_slot, BN); tc::tmem_relinquish(); }
    for (int i = threadIdx.x; i < BN; i += LGL_THREADS(BN)) {
        s_par[i] = bias[n0 + i];
        if (MODE == LG_FC1) { s_par[BN + i] = ln_g[i]; s_par[2 * BN + i] = ln_b[i]; }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr int NCHUNK = KIN / 64;

    if (warp == 0) {
        if (lane == 0) {
            for (int c = 0; c < NCHUNK; ++c) {
                const int s = c % LGL_STAGES;
                if (c >= LGL_STAGES && !tc::mbar_wait(&empty[s], ((c / LGL_STAGES) & 1) ^ 1, err, 501)) break;
                uint8_t* sa = smem + s * Cfg::STAGE;
                uint8_t* sb = sa + Cfg::A_BYTES;
                tc::mbar_arrive_expect_tx(&full[s], Cfg::STAGE);
                tc::tma_load_3d(sa, &map_a, &full[s], c * 64, r0, slot);
                for (int nb = 0; nb < BN / 256; ++nb) tc::tma_load_3d(sb + nb * 256 * 128, &map_w, &full[s], c * 64, n0 + nb * 256, 0);
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = tc::make_idesc_bf16(128, 256);
        for (int c = 0; c < NCHUNK; ++c) {
            const int s = c % LGL_STAGES;
            if (!tc::mbar_wait(&full[s], (c / LGL_STAGES) & 1, err, 502)) break;
            tc::tc_fence_after();
            if (tc::elect_one()) {
                const uint32_t a_addr = tc::smem_u32(smem + s * Cfg::STAGE);
                const uint64_t da0 = tc::make_smem_desc_sw128(a_addr, 1024);
                const uint64_t db0 = tc::make_smem_desc_sw128(a_addr + Cfg::A_BYTES, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int nb = 0; nb < BN / 256; ++nb)
                        tc::umma_bf16(tmem_base + (uint32_t)(nb * 256), da0 + (uint64_t)((k * 32) >> 4),
                                      db0 + (uint64_t)((nb * 256 * 128 + k * 32) >> 4), idesc, (c | k) ? 1u : 0u);
                tc::umma_commit(&empty[s]);
                if (c == NCHUNK - 1) tc::umma_commit(acc_full);
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        const int qd = warp & 3, m = qd * 32 + lane;
        const bool ok = tc::mbar_wait(acc_full, 0, err, 503);
        tc::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16);
        if (ok) {
            if (MODE == LG_FC1) {
                // LayerNorm over the 512 outputs of this token: mean, then variance about the mean, then normalise +
                // exact GELU -> bf16 hidden activation.  Three sweeps over the TMEM-resident row; the two column groups
                // exchange their partial statistics through shared memory.
                const int grp = (warp - 4) >> 2, cb = grp * (BN / 2), ce = cb + BN / 2;
                float sum = 0.f;
#pragma unroll 1
                for (int c0 = cb; c0 < ce; c0 += 32) {
                    uint32_t r[32];
                    tc::tmem_ld32(taddr + c0, r);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) sum += __uint_as_float(r[i]) + s_par[c0 + i];
                }
                s_stat[grp * 128 + m] = sum;
                tc::named_bar_sync(1, 256);
                const float mean = (s_stat[m] + s_stat[128 + m]) * (1.0f / BN);
                float var = 0.f;
#pragma unroll 1
                for (int c0 = cb; c0 < ce; c0 += 32) {
                    uint32_t r[32];
                    tc::tmem_ld32(taddr + c0, r);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float d = __uint_as_float(r[i]) + s_par[c0 + i] - mean;
                        var = fmaf(d, d, var);
                    }
                }
                s_stat[256 + grp * 128 + m] = var;
                tc::named_bar_sync(1, 256);
                const float rstd = rsqrtf((s_stat[256 + m] + s_stat[384 + m]) * (1.0f / BN) + 1e-5f);
                // write-out through this warp's 4 KB staging tile: whole 128-byte row segments per 8 lanes
                uint8_t* stg = s_stage + (warp - 4) * 4096;
                const int wrow0 = r0 + qd * 32;
                const size_t tok0 = (size_t)slot * k_cap + wrow0;
#pragma unroll 1
                for (int c0 = cb; c0 < ce; c0 += 32) {
                    uint32_t r[32];
                    tc::tmem_ld32(taddr + c0, r);
                    tc::tmem_ld_wait();
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 bb = *reinterpret_cast<const float4*>(&s_par[c0 + i]);
                        const float4 gg = *reinterpret_cast<const float4*>(&s_par[BN + c0 + i]);
                        const float4 be = *reinterpret_cast<const float4*>(&s_par[2 * BN + c0 + i]);
                        const float bv[4] = {bb.x, bb.y, bb.z, bb.w}, gv[4] = {gg.x, gg.y, gg.z, gg.w}, ev[4] = {be.x, be.y, be.z, be.w};
                        float v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float y = (__uint_as_float(r[i + u]) + bv[u] - mean) * rstd * gv[u] + ev[u];
                            v[u] = 0.5f * y * (1.0f + lg_erf_fast(y * 0.70710678118654752f));
                        }
                        pk[i >> 1] = tc::pack_bf16x2(v[0], v[1]);
                        pk[(i >> 1) + 1] = tc::pack_bf16x2(v[2], v[3]);
                    }
                    const int half = (c0 >> 5) & 1;
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 * half + g) ^ (lane & 7)) << 4)) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
                    if (half) {   // a 64-column panel (128 B per row) is complete
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int q = lane + 32 * i, rr = q >> 3, j = q & 7;
                            if (wrow0 + rr < n)
                                *reinterpret_cast<uint4*>(out0 + (tok0 + rr) * LG_HID + c0 - 32 + j * 8) =
                                    *reinterpret_cast<const uint4*>(stg + rr * 128 + ((j ^ (rr & 7)) << 4));
                        }
                        __syncwarp();
                    }
                }
            }
        }
        tc::tc_fence_before();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, BN); }
}

// ---- persistent, weight-resident linear layers (QKV / OUT / FC2) ---------------------------------------------------
// The one-tile-per-CTA kernel above re-streams its whole weight slice (128 KB) from L2 for every 128 tokens, which
// made the N = 256 layers L2-bound.  Here a CTA loads its weight slice ONCE (KIN x BN bf16 = 128 KB, all K chunks
// resident in the 128B-swizzle layout), then loops over token tiles: only the A tile (16 KB per K chunk, 4-stage TMA
// ring) moves, two TMEM accumulator stages overlap the epilogue of tile i with the MMAs of tile i + 1, and eight
// epilogue warps (two column groups) drain the accumulators.  grid = (CTAs per weight slice, slices).
#define LGP_STAGES 4
#define LGP_THREADS 384
template <int KIN, int BN> struct LgpCfg {
    static constexpr int W_BYTES = KIN * BN * 2;
    static constexpr int A_BYTES = 128 * 128;
    static constexpr int SMEM = 1024 + W_BYTES + LGP_STAGES * A_BYTES + 256 + BN * 4 + 8 * 4096;   // + one 4 KB staging tile per epilogue warp
};

template <int MODE, int KIN, int BN>
__global__ void __launch_bounds__(LGP_THREADS, 1)
lg_linear_persist_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const int* __restrict__ kp_count,
                         LgPairs pr, int k_cap, int row_tiles, const float* __restrict__ bias, int rotary, const float* __restrict__ cs,
                         bf16* __restrict__ out0, bf16* __restrict__ out1, bf16* __restrict__ out2, float* __restrict__ xres, int* err) {
    typedef LgpCfg<KIN, BN> Cfg;
    constexpr int NCHUNK = KIN / 64;
    const int part = blockIdx.y, n0 = part * BN;     // weight slice = output columns [n0, n0 + BN)
    const int total_tiles = 2 * pr.pairs * row_tiles;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW = smem;                              // NCHUNK x [BN rows x 128 B]
    uint8_t* sA = smem + Cfg::W_BYTES;               // LGP_STAGES x [128 rows x 128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + LGP_STAGES * Cfg::A_BYTES);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [BN] bias of this slice
    uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_bias + BN);                          // 8 x 4 KB, one per epilogue warp
    uint64_t* w_full = bars;                 // 1
    uint64_t* a_full = bars + 1;             // [4]
    uint64_t* a_empty = bars + 5;            // [4]
    uint64_t* t_full = bars + 9;             // [2]
    uint64_t* t_empty = bars + 11;           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&map_a);
        tc::tma_prefetch_desc(&map_w);
        tc::mbar_init(w_full, 1);
        for (int s = 0; s < LGP_STAGES; ++s) { tc::mbar_init(&a_full[s], 1); tc::mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&t_full[s], 1); tc::mbar_init(&t_empty[s], 8); }
        tc::fence_barrier_init();
    }
    if (warp == 2) { tc::tmem_alloc(tmem_slot, 2 * BN); tc::tmem_relinquish(); }
    for (int i = threadIdx.x; i < BN; i += LGP_THREADS) s_bias[i] = bias[n0 + i];
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // every role walks the same tile sequence and skips the same empty tiles
    auto tile_info = [&](int t, int& slot, int& r0) -> bool {
        int partner;
        lg_slots(pr, t / row_tiles, slot, partner);
        r0 = (t % row_tiles) * 128;
        return r0 < max(kp_count[slot], 0) && max(kp_count[partner], 0) > 0;
    };

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(w_full, Cfg::W_BYTES);
            for (int c = 0; c < NCHUNK; ++c) tc::tma_load_3d(sW + c * BN * 128, &map_w, w_full, c * 64, n0, 0);
            int g = 0;
            bool ok = true;
            for (int t = blockIdx.x; ok && t < total_tiles; t += gridDim.x) {
                int slot, r0;
                if (!tile_info(t, slot, r0)) continue;
                for (int c = 0; c < NCHUNK; ++c, ++g) {
                    const int s = g % LGP_STAGES;
                    if (g >= LGP_STAGES && !tc::mbar_wait(&a_empty[s], ((g / LGP_STAGES) & 1) ^ 1, err, 531)) { ok = false; break; }
                    tc::mbar_arrive_expect_tx(&a_full[s], Cfg::A_BYTES);
                    tc::tma_load_3d(sA + s * Cfg::A_BYTES, &map_a, &a_full[s], c * 64, r0, slot);
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = tc::make_idesc_bf16(128, BN);
        bool ok = tc::mbar_wait(w_full, 0, err, 532);
        const uint64_t dw0 = tc::make_smem_desc_sw128(tc::smem_u32(sW), 1024);
        const uint64_t da00 = tc::make_smem_desc_sw128(tc::smem_u32(sA), 1024);
        int g = 0, ti = 0;
        for (int t = blockIdx.x; ok && t < total_tiles; t += gridDim.x) {
            int slot, r0;
            if (!tile_info(t, slot, r0)) continue;
            const int as = ti & 1;
            if (ti >= 2 && !tc::mbar_wait(&t_empty[as], ((ti >> 1) & 1) ^ 1, err, 533)) break;
            for (int c = 0; c < NCHUNK; ++c, ++g) {
                const int s = g % LGP_STAGES;
                if (!tc::mbar_wait(&a_full[s], (g / LGP_STAGES) & 1, err, 534)) { ok = false; break; }
                tc::tc_fence_after();
                if (tc::elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc::umma_bf16(tmem_base + (uint32_t)(as * BN), da00 + (uint64_t)((s * Cfg::A_BYTES + k * 32) >> 4),
                                      dw0 + (uint64_t)((c * BN * 128 + k * 32) >> 4), idesc, (c | k) ? 1u : 0u);
                    tc::umma_commit(&a_empty[s]);
                    if (c == NCHUNK - 1) tc::umma_commit(&t_full[as]);
                }
                __syncwarp();
            }
            ++ti;
        }
    } else if (warp >= 4) {
        // Epilogue: thread <-> accumulator row (TMEM lane).  The read-out is latency-bound (two warps per scheduler),
        // so the next 32-column TMEM load and this group's global operands (residual / rotary table) are in flight
        // while the current group is converted and stored; the bias comes from shared memory.
        const int qd = warp & 3, m = qd * 32 + lane, grp = (warp - 4) >> 2;
        constexpr int NG = BN / 2 / 32;                       // 32-column groups per thread and tile
        uint8_t* stg = s_stage + (warp - 4) * 4096;
        const int cb = grp * (BN / 2);                        // this warpgroup's first column of the slice
        int ti = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            int slot, r0;
            if (!tile_info(t, slot, r0)) continue;
            const int as = ti & 1;
            if (!tc::mbar_wait(&t_full[as], (ti >> 1) & 1, err, 535)) break;
            tc::tc_fence_after();
            const int row = r0 + m, n_rows = max(kp_count[slot], 0);
            const bool valid = row < n_rows;
            const int wrow0 = r0 + qd * 32;                              // first row of this warp's 32-row band
            const size_t tok0 = (size_t)slot * k_cap + wrow0;
            const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(as * BN + cb);
            uint32_t r[2][32];
            tc::tmem_ld32(taddr, r[0]);
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const int c0 = cb + 32 * g, gc = n0 + c0;      // column in the slice / global output column
                // global operands of this group first, so their latency overlaps the TMEM wait
                float4 xv[8];
                float cv[16], sv[16];
                if (MODE == LG_FC2) {
                    // residual in the write-out mapping (8 lanes per 128-byte row segment)
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int q = lane + 32 * i, rr = q >> 3, j = q & 7;
                        xv[i] = (wrow0 + rr < n_rows) ? *reinterpret_cast<const float4*>(xres + (tok0 + rr) * LG_DIM + gc + j * 4)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                if (MODE == LG_QKV && part < 2 && rotary && valid) {
                    // columns c0..c0+31 of a 64-wide head: pairs (2i, 2i+1) share angle (c0 & 63)/2 + i
                    const float* ct = cs + ((size_t)slot * 64 + ((c0 & 63) >> 1)) * k_cap + row;
#pragma unroll
                    for (int i = 0; i < 16; ++i) { cv[i] = __ldg(ct + (size_t)i * k_cap); sv[i] = __ldg(ct + (size_t)(32 + i) * k_cap); }
                }
                tc::tmem_ld_wait();
                if (g + 1 < NG) tc::tmem_ld32(taddr + 32 * (g + 1), r[(g + 1) & 1]);
                float v[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 bb = *reinterpret_cast<const float4*>(&s_bias[c0 + 4 * j]);
                    v[4 * j] = __uint_as_float(r[g & 1][4 * j]) + bb.x; v[4 * j + 1] = __uint_as_float(r[g & 1][4 * j + 1]) + bb.y;
                    v[4 * j + 2] = __uint_as_float(r[g & 1][4 * j + 2]) + bb.z; v[4 * j + 3] = __uint_as_float(r[g & 1][4 * j + 3]) + bb.w;
                }
                if (MODE == LG_QKV && part == 2) {
                    // V transposed [slot][channel][token]: consecutive lanes = consecutive tokens, one line per store
                    if (valid) {
                        bf16* dst = out2 + ((size_t)slot * LG_DIM + c0) * k_cap + row;
#pragma unroll
                        for (int i = 0; i < 32; ++i) dst[(size_t)i * k_cap] = __float2bfloat16_rn(v[i]);
                    }
                    continue;
                }
                // Row-major outputs go through this warp's 4 KB staging tile ([32 rows][128 B], 16-byte pieces XOR-swizzled
                // by row) so that one store instruction covers four whole 128-byte row segments instead of 32 scattered
                // 16-byte pieces (the L1 tag stage takes one cycle per distinct line).
                if (MODE == LG_FC2) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int q = lane + 32 * i, rr = q >> 3, j = q & 7;
                        const float4 a = *reinterpret_cast<const float4*>(stg + rr * 128 + ((j ^ (rr & 7)) << 4));
                        if (wrow0 + rr < n_rows) {
                            float4 x = xv[i];
                            x.x = __fadd_rn(x.x, a.x); x.y = __fadd_rn(x.y, a.y); x.z = __fadd_rn(x.z, a.z); x.w = __fadd_rn(x.w, a.w);
                            *reinterpret_cast<float4*>(xres + (tok0 + rr) * LG_DIM + gc + j * 4) = x;
                            *reinterpret_cast<uint2*>(out0 + (tok0 + rr) * LG_HID + gc + j * 4) = make_uint2(tc::pack_bf16x2(x.x, x.y), tc::pack_bf16x2(x.z, x.w));
                        }
                    }
                    __syncwarp();
                } else {
                    if (MODE == LG_QKV && rotary) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float e = v[2 * i], o = v[2 * i + 1];
                            v[2 * i] = __fadd_rn(__fmul_rn(e, cv[i]), __fmul_rn(-o, sv[i]));
                            v[2 * i + 1] = __fadd_rn(__fmul_rn(o, cv[i]), __fmul_rn(e, sv[i]));
                        }
                    }
                    const float sc = (MODE == LG_QKV && part == 0) ? 0.125f : 1.0f;   // 1/sqrt(head_dim) folded into q (exact)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 * (g & 1) + j) ^ (lane & 7)) << 4)) =
                            make_uint4(tc::pack_bf16x2(v[8 * j] * sc, v[8 * j + 1] * sc), tc::pack_bf16x2(v[8 * j + 2] * sc, v[8 * j + 3] * sc),
                                       tc::pack_bf16x2(v[8 * j + 4] * sc, v[8 * j + 5] * sc), tc::pack_bf16x2(v[8 * j + 6] * sc, v[8 * j + 7] * sc));
                    if (g & 1) {   // a 64-column panel (128 B per row) is complete
                        __syncwarp();
                        const int pc0 = c0 - 32;   // first column of the panel in the slice
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int q = lane + 32 * i, rr = q >> 3, j = q & 7;
                            if (wrow0 + rr < n_rows) {
                                bf16* dst = MODE == LG_OUT ? out0 + (tok0 + rr) * LG_HID + LG_DIM + n0 + pc0 + j * 8
                                                           : (part == 0 ? out0 : out1) + (tok0 + rr) * LG_DIM + pc0 + j * 8;
                                *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(stg + rr * 128 + ((j ^ (rr & 7)) << 4));
                            }
                        }
                        __syncwarp();
                    }
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&t_empty[as]);
            ++ti;
        }
        tc::tc_fence_before();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 2 * BN); }
}

// ---- attention ------------------------------------------------------------------------------------------------
#define LGA_Q_BYTES (128 * 128)
#define LGA_K_BYTES (128 * 128)
#define LGA_V_BYTES (2 * 64 * 128)       // two 64-key chunks of [64 dims][128 B]
#define LGA_P_BYTES (2 * 128 * 128)      // two 64-key chunks of [128 queries][128 B]
#define LGA_SMEM (1024 + LGA_Q_BYTES + 2 * LGA_K_BYTES + 2 * LGA_V_BYTES + 2 * LGA_P_BYTES + 256 + 4 * 128 * 4)
#define LGA_THREADS 384   // warp 0 TMA, 1 MMA, 2 TMEM alloc, 4-11 softmax

__global__ void __launch_bounds__(LGA_THREADS, 1)
lg_attn_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_vt,
               const int* __restrict__ kp_count, LgPairs pr, int k_cap, int cross, bf16* __restrict__ att, int* err) {
    int slot, partner;
    lg_slots(pr, blockIdx.z, slot, partner);
    const int head = blockIdx.y;
    const int n_q = max(kp_count[slot], 0);
    const int slot_kv = cross ? partner : slot;
    const int n_kv = max(kp_count[slot_kv], 0);
    const int r0 = blockIdx.x * 128;
    if (r0 >= n_q || max(kp_count[partner], 0) == 0) return;
    const int nt = (n_kv + 127) / 128;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + LGA_Q_BYTES;
    uint8_t* sV = sK + 2 * LGA_K_BYTES;
    uint8_t* sP = sV + 2 * LGA_V_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * LGA_P_BYTES);
    uint64_t* q_full = bars;            // 1
    uint64_t* k_full = bars + 1;        // [2]
    uint64_t* k_empty = bars + 3;       // [2]
    uint64_t* v_full = bars + 5;        // [2]
    uint64_t* v_empty = bars + 7;       // [2]
    uint64_t* s_full = bars + 9;        // [2]
    uint64_t* s_empty = bars + 11;      // [2]
    uint64_t* p_full = bars + 13;       // [2]
    uint64_t* p_empty = bars + 15;      // [2]
    uint64_t* o_full = bars + 17;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    float* s_mx = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2][128] per-group row max
    float* s_sm = s_mx + 256;                                                          // [2][128] per-group row sum
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&map_q);
        tc::tma_prefetch_desc(&map_k);
        tc::tma_prefetch_desc(&map_vt);
        tc::mbar_init(q_full, 1);
        tc::mbar_init(o_full, 1);
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&k_full[s], 1); tc::mbar_init(&k_empty[s], 1);
            tc::mbar_init(&v_full[s], 1); tc::mbar_init(&v_empty[s], 1);
            tc::mbar_init(&s_full[s], 1); tc::mbar_init(&s_empty[s], 8);
            tc::mbar_init(&p_full[s], 8); tc::mbar_init(&p_empty[s], 1);
        }
        tc::fence_barrier_init();
    }
    if (warp == 2) { tc::tmem_alloc(tmem_slot, 512); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int total = 2 * nt;   // pass 0 tiles then pass 1 tiles

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(q_full, LGA_Q_BYTES);
            tc::tma_load_3d(sQ, &map_q, q_full, head * LG_HD, r0, slot);
            for (int it = 0; it < total; ++it) {
                const int j = it >= nt ? it - nt : it, s = it & 1;
                if (it >= 2 && !tc::mbar_wait_fast(&k_empty[s], ((it >> 1) & 1) ^ 1, err, 511)) break;
                tc::mbar_arrive_expect_tx(&k_full[s], LGA_K_BYTES);
                tc::tma_load_3d(sK + s * LGA_K_BYTES, &map_k, &k_full[s], head * LG_HD, j * 128, slot_kv);
                if (it >= nt) {
                    const int sv = j & 1;
                    if (j >= 2 && !tc::mbar_wait_fast(&v_empty[sv], ((j >> 1) & 1) ^ 1, err, 512)) break;
                    tc::mbar_arrive_expect_tx(&v_full[sv], LGA_V_BYTES);
                    tc::tma_load_3d(sV + sv * LGA_V_BYTES, &map_vt, &v_full[sv], j * 128, 0, slot_kv * LG_HEADS + head);
                    tc::tma_load_3d(sV + sv * LGA_V_BYTES + 64 * 128, &map_vt, &v_full[sv], j * 128 + 64, 0, slot_kv * LG_HEADS + head);
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc_s = tc::make_idesc_bf16(128, 128), idesc_o = tc::make_idesc_bf16(128, 64);
        bool ok = tc::mbar_wait_fast(q_full, 0, err, 513);
        const uint64_t dq = tc::make_smem_desc_sw128(tc::smem_u32(sQ), 1024);
        // software pipeline: S(it) is issued before P V(it - 1), so the softmax warps always have a tile to chew on
        for (int it = 0; ok && it <= total; ++it) {
            if (it < total) {
                const int s = it & 1;
                if (!tc::mbar_wait_fast(&k_full[s], (it >> 1) & 1, err, 514)) break;
                if (it >= 2 && !tc::mbar_wait_fast(&s_empty[s], ((it >> 1) & 1) ^ 1, err, 515)) break;
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint64_t dk = tc::make_smem_desc_sw128(tc::smem_u32(sK + s * LGA_K_BYTES), 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc::umma_bf16(tmem_base + (uint32_t)(s * 128), dq + (uint64_t)((k * 32) >> 4), dk + (uint64_t)((k * 32) >> 4), idesc_s, k ? 1u : 0u);
                    tc::umma_commit(&k_empty[s]);
                    tc::umma_commit(&s_full[s]);
                }
                __syncwarp();
            }
            const int pit = it - 1;
            if (pit >= nt) {
                const int j = pit - nt, jb = j & 1;
                if (!tc::mbar_wait_fast(&p_full[jb], (j >> 1) & 1, err, 516)) break;
                if (!tc::mbar_wait_fast(&v_full[jb], (j >> 1) & 1, err, 517)) break;
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint64_t dp = tc::make_smem_desc_sw128(tc::smem_u32(sP + jb * LGA_P_BYTES), 1024);
                    const uint64_t dv = tc::make_smem_desc_sw128(tc::smem_u32(sV + jb * LGA_V_BYTES), 1024);
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc::umma_bf16(tmem_base + 256u, dp + (uint64_t)((c * 128 * 128 + k * 32) >> 4),
                                          dv + (uint64_t)((c * 64 * 128 + k * 32) >> 4), idesc_o, (j | c | k) ? 1u : 0u);
                    tc::umma_commit(&v_empty[jb]);
                    tc::umma_commit(&p_empty[jb]);
                    if (j == nt - 1) tc::umma_commit(o_full);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===== softmax: 8 warps; thread <-> query row (TMEM lane), group `grp` owns key columns [64 grp, 64 grp + 64)
        // of every 128-key tile, i.e. one 64-key chunk of the P operand =====
        const int qd = warp & 3, m = qd * 32 + lane, row = r0 + m;
        const int grp = (warp - 4) >> 2;
        const float LOG2E = 1.4426950408889634f;
        float run_max = -INFINITY, run_sum = 0.f, neg_ml2 = 0.f;
        bool ok = true;
        for (int it = 0; ok && it < total; ++it) {
            const int pass = it >= nt, j = pass ? it - nt : it, s = it & 1, c0 = j * 128 + grp * 64;
            if (it == nt) {
                // row max over all keys = max of the two groups' halves
                s_mx[grp * 128 + m] = run_max;
                tc::named_bar_sync(1, 256);
                neg_ml2 = -fmaxf(s_mx[m], s_mx[128 + m]) * LOG2E;
            }
            if (!tc::mbar_wait_fast(&s_full[s], (it >> 1) & 1, err, 518)) { ok = false; break; }
            const int jb = j & 1;
            if (pass && j >= 2 && !tc::mbar_wait_fast(&p_empty[jb], ((j >> 1) & 1) ^ 1, err, 519)) { ok = false; break; }
            tc::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(s * 128 + grp * 64);
            uint32_t v[64];
            tc::tmem_ld32(taddr, v);
            tc::tmem_ld32(taddr + 32, v + 32);
            tc::tmem_ld_wait();
            const bool full = c0 + 64 <= n_kv;
            if (!pass) {
                if (full) {
#pragma unroll
                    for (int i = 0; i < 64; ++i) run_max = fmaxf(run_max, __uint_as_float(v[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 64; ++i)
                        if (c0 + i < n_kv) run_max = fmaxf(run_max, __uint_as_float(v[i]));
                }
            } else {
                // P = exp(S - rowmax) (unnormalised, <= 1) as the bf16 A operand of the second MMA; the row sum of the
                // unrounded values normalises O at the end
                float p[64];
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    float e;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(__uint_as_float(v[i]), LOG2E, neg_ml2)));
                    p[i] = e;
                }
                if (!full) {
#pragma unroll
                    for (int i = 0; i < 64; ++i)
                        if (c0 + i >= n_kv) p[i] = 0.f;
                }
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;   // four chains instead of one 64-long dependent FADD chain
#pragma unroll
                for (int i = 0; i < 64; i += 4) { s0 += p[i]; s1 += p[i + 1]; s2 += p[i + 2]; s3 += p[i + 3]; }
                run_sum += (s0 + s1) + (s2 + s3);
                uint8_t* prow = sP + jb * LGA_P_BYTES + grp * (128 * 128) + m * 128;
#pragma unroll
                for (int g = 0; g < 8; ++g)
                    *reinterpret_cast<uint4*>(prow + ((g ^ (m & 7)) << 4)) =
                        make_uint4(tc::pack_bf16x2(p[8 * g], p[8 * g + 1]), tc::pack_bf16x2(p[8 * g + 2], p[8 * g + 3]),
                                   tc::pack_bf16x2(p[8 * g + 4], p[8 * g + 5]), tc::pack_bf16x2(p[8 * g + 6], p[8 * g + 7]));
            }
            tc::tc_fence_before();
            if (pass) tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tc::mbar_arrive(&s_empty[s]);
                if (pass) tc::mbar_arrive(&p_full[jb]);
            }
        }
        s_sm[grp * 128 + m] = run_sum;
        tc::named_bar_sync(2, 256);
        const float inv = __fdiv_rn(1.0f, s_sm[m] + s_sm[128 + m]);
        if (ok && tc::mbar_wait_fast(o_full, 0, err, 520)) {
            tc::tc_fence_after();
            // O is 64 columns wide: group g converts columns [32 g, 32 g + 32)
            const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + 256u + (uint32_t)(grp * 32);
            uint32_t v[32];
            tc::tmem_ld32(taddr, v);
            tc::tmem_ld_wait();
            if (row < n_q) {
                bf16* dst = att + ((size_t)slot * k_cap + row) * LG_DIM + head * LG_HD + grp * 32;
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    reinterpret_cast<uint4*>(dst)[g] = make_uint4(
                        tc::pack_bf16x2(__uint_as_float(v[8 * g]) * inv, __uint_as_float(v[8 * g + 1]) * inv),
                        tc::pack_bf16x2(__uint_as_float(v[8 * g + 2]) * inv, __uint_as_float(v[8 * g + 3]) * inv),
                        tc::pack_bf16x2(__uint_as_float(v[8 * g + 4]) * inv, __uint_as_float(v[8 * g + 5]) * inv),
                        tc::pack_bf16x2(__uint_as_float(v[8 * g + 6]) * inv, __uint_as_float(v[8 * g + 7]) * inv));
            }
        }
        tc::tc_fence_before();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}

// ---- host ------------------------------------------------------------------------------------------------------
static inline uint16_t lg_bf16_bits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

void gnb_lightglue_free(gnb_ctx* ctx) {
    LgState* st = lg_state(ctx);
    if (!st) return;
    void* ptrs[] = {st->wblob, st->xo, st->q, st->k, st->vt, st->att, st->h, st->cs};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    delete st;
    ctx->lg_state = nullptr;
}

extern "C" int gnb_matcher_layers(const gnb_ctx* ctx) { return (ctx && ctx->lg_state) ? lg_state(const_cast<gnb_ctx*>(ctx))->n_layers : 0; }

static int lg_load(gnb_ctx* ctx, const void* blob, size_t nbytes);

extern "C" int gnb_set_matcher_layers(gnb_ctx* ctx, const void* blob, size_t nbytes) {
    if (!ctx) return GNB_E_INVALID;
    const int rc = lg_load(ctx, blob, nbytes);
    if (rc != GNB_OK) gnb_lightglue_free(ctx);   // never leave a half-built layer state behind: the matcher falls back to the head
    return rc;
}

static int lg_load(gnb_ctx* ctx, const void* blob, size_t nbytes) {
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    GNB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    gnb_lightglue_free(ctx);
    gnb_cache_clear(ctx);                      // cached raster features are specific to the matcher configuration
    if (!blob || nbytes == 0) return GNB_OK;   // back to the head-only matcher
    if (ctx->cfg.match_impl != 0) { GNB_SET_ERR(ctx, "transformer layers need the tcgen05 matcher (match_impl = 0)"); return GNB_E_INVALID; }
    const int kc = ctx->cfg.max_keypoints;
    if (kc % 8) { GNB_SET_ERR(ctx, "transformer layers need max_keypoints to be a multiple of 8 (got %d)", kc); return GNB_E_INVALID; }
    uint32_t hdr[4];
    if (nbytes < 16) { GNB_SET_ERR(ctx, "layer blob too small"); return GNB_E_INVALID; }
    memcpy(hdr, blob, 16);
    const size_t per_block = 4 * (256 * 256 + 256) + 512 * 512 + 512 + 512 + 512 + 256 * 512 + 256;
    const int n_layers = (int)hdr[2];
    const size_t expect = 64 + (size_t)n_layers * 2 * per_block;
    if (memcmp(blob, "GNBL", 4) != 0 || hdr[1] != 1 || n_layers < 1 || n_layers > 64 || hdr[3] != expect || nbytes != 16 + 4 * expect) {
        GNB_SET_ERR(ctx, "bad layer blob (magic/version/size)");
        return GNB_E_INVALID;
    }
    const float* fl = reinterpret_cast<const float*>(static_cast<const uint8_t*>(blob) + 16);
    LgState* st = new LgState();
    st->wblob = nullptr; st->xo = st->q = st->k = st->vt = st->att = st->h = nullptr; st->cs = nullptr;
    st->n_layers = n_layers;
    ctx->lg_state = st;
    memcpy(st->pos_w, fl, 64 * sizeof(float));
    fl += 64;
    // device image of the weights: per block bf16 matrices then fp32 vectors, every piece 256-byte aligned
    const size_t mat_elems = 768 * 256 + 256 * 256 + 512 * 512 + 256 * 512;   // bf16
    const size_t vec_elems = 768 + 256 + 512 + 512 + 512 + 256;               // f32
    const size_t blk_bytes = ((mat_elems * 2 + 255) & ~(size_t)255) + ((vec_elems * 4 + 255) & ~(size_t)255);
    const size_t total = blk_bytes * 2 * n_layers;
    std::vector<uint8_t> img(total, 0);
    GNB_CUDA(ctx, cudaMalloc(&st->wblob, total));
    st->blocks.resize(2 * n_layers);
    for (int b = 0; b < 2 * n_layers; ++b) {
        uint8_t* base = img.data() + blk_bytes * b;
        uint16_t* mats = reinterpret_cast<uint16_t*>(base);
        float* vecs = reinterpret_cast<float*>(base + ((mat_elems * 2 + 255) & ~(size_t)255));
        // blob order: q.w q.b k.w k.b v.w v.b o.w o.b fc1.w fc1.b ln.w ln.b fc2.w fc2.b
        const float* p = fl + per_block * b;
        uint16_t* wqkv = mats; uint16_t* wo = wqkv + 768 * 256; uint16_t* w1 = wo + 256 * 256; uint16_t* w2 = w1 + 512 * 512;
        float* bqkv = vecs; float* bo = bqkv + 768; float* b1 = bo + 256; float* lng = b1 + 512; float* lnb = lng + 512; float* b2 = lnb + 512;
        for (int part = 0; part < 3; ++part) {
            for (int i = 0; i < 256 * 256; ++i) wqkv[part * 256 * 256 + i] = lg_bf16_bits(p[i]);
            p += 256 * 256;
            memcpy(bqkv + part * 256, p, 256 * 4);
            p += 256;
        }
        for (int i = 0; i < 256 * 256; ++i) wo[i] = lg_bf16_bits(p[i]);
        p += 256 * 256;
        memcpy(bo, p, 256 * 4); p += 256;
        for (int i = 0; i < 512 * 512; ++i) w1[i] = lg_bf16_bits(p[i]);
        p += 512 * 512;
        memcpy(b1, p, 512 * 4); p += 512;
        memcpy(lng, p, 512 * 4); p += 512;
        memcpy(lnb, p, 512 * 4); p += 512;
        for (int i = 0; i < 256 * 512; ++i) w2[i] = lg_bf16_bits(p[i]);
        p += 256 * 512;
        memcpy(b2, p, 256 * 4);
        uint8_t* dbase = static_cast<uint8_t*>(st->wblob) + blk_bytes * b;
        LgBlockW& w = st->blocks[b];
        w.wqkv = reinterpret_cast<bf16*>(dbase); w.wo = w.wqkv + 768 * 256; w.w1 = w.wo + 256 * 256; w.w2 = w.w1 + 512 * 512;
        float* dv = reinterpret_cast<float*>(dbase + ((mat_elems * 2 + 255) & ~(size_t)255));
        w.bqkv = dv; w.bo = dv + 768; w.b1 = w.bo + 256; w.lng = w.b1 + 512; w.lnb = w.lng + 512; w.b2 = w.lnb + 512;
        int rc = 0;
        const uint32_t box[3] = {64, 256, 1};
        { const uint64_t d[3] = {256, 768, 1}, s[2] = {512, 768 * 512}; rc |= gnb_make_tmap_bf16(ctx, &w.m_qkv, w.wqkv, 3, d, s, box); }
        { const uint64_t d[3] = {256, 256, 1}, s[2] = {512, 256 * 512}; rc |= gnb_make_tmap_bf16(ctx, &w.m_o, w.wo, 3, d, s, box); }
        { const uint64_t d[3] = {512, 512, 1}, s[2] = {1024, 512 * 1024}; rc |= gnb_make_tmap_bf16(ctx, &w.m_w1, w.w1, 3, d, s, box); }
        { const uint32_t b2[3] = {64, 128, 1};
          const uint64_t d[3] = {512, 256, 1}, s[2] = {1024, 256 * 1024}; rc |= gnb_make_tmap_bf16(ctx, &w.m_w2, w.w2, 3, d, s, b2); }
        if (rc) { gnb_lightglue_free(ctx); return GNB_E_CUDA; }
    }
    GNB_CUDA(ctx, cudaMemcpy(st->wblob, img.data(), total, cudaMemcpyHostToDevice));
    // activations
    const size_t slots = (size_t)ctx->kp_slots, tok = slots * kc;
    GNB_CUDA(ctx, cudaMalloc(&st->xo, tok * LG_HID * 2));
    GNB_CUDA(ctx, cudaMalloc(&st->q, tok * LG_DIM * 2));
    GNB_CUDA(ctx, cudaMalloc(&st->k, tok * LG_DIM * 2));
    GNB_CUDA(ctx, cudaMalloc(&st->vt, tok * LG_DIM * 2));
    GNB_CUDA(ctx, cudaMalloc(&st->att, tok * LG_DIM * 2));
    GNB_CUDA(ctx, cudaMalloc(&st->h, tok * LG_HID * 2));
    GNB_CUDA(ctx, cudaMalloc(&st->cs, tok * 64 * 4));
    GNB_CUDA(ctx, cudaMemset(st->xo, 0, tok * LG_HID * 2));
    GNB_CUDA(ctx, cudaMemset(st->q, 0, tok * LG_DIM * 2));
    GNB_CUDA(ctx, cudaMemset(st->k, 0, tok * LG_DIM * 2));
    GNB_CUDA(ctx, cudaMemset(st->vt, 0, tok * LG_DIM * 2));
    GNB_CUDA(ctx, cudaMemset(st->att, 0, tok * LG_DIM * 2));
    GNB_CUDA(ctx, cudaMemset(st->h, 0, tok * LG_HID * 2));
    int rc = 0;
    const uint32_t box[3] = {64, 128, 1};
    { const uint64_t d[3] = {256, (uint64_t)kc, slots}, s[2] = {LG_HID * 2, (uint64_t)kc * LG_HID * 2}; rc |= gnb_make_tmap_bf16(ctx, &st->m_x, st->xo, 3, d, s, box); }
    { const uint64_t d[3] = {512, (uint64_t)kc, slots}, s[2] = {LG_HID * 2, (uint64_t)kc * LG_HID * 2}; rc |= gnb_make_tmap_bf16(ctx, &st->m_xo, st->xo, 3, d, s, box); }
    { const uint64_t d[3] = {256, (uint64_t)kc, slots}, s[2] = {LG_DIM * 2, (uint64_t)kc * LG_DIM * 2}; rc |= gnb_make_tmap_bf16(ctx, &st->m_att, st->att, 3, d, s, box); }
    { const uint64_t d[3] = {512, (uint64_t)kc, slots}, s[2] = {LG_HID * 2, (uint64_t)kc * LG_HID * 2}; rc |= gnb_make_tmap_bf16(ctx, &st->m_h, st->h, 3, d, s, box); }
    { const uint64_t d[3] = {256, (uint64_t)kc, slots}, s[2] = {LG_DIM * 2, (uint64_t)kc * LG_DIM * 2}; rc |= gnb_make_tmap_bf16(ctx, &st->m_q, st->q, 3, d, s, box); }
    { const uint64_t d[3] = {256, (uint64_t)kc, slots}, s[2] = {LG_DIM * 2, (uint64_t)kc * LG_DIM * 2}; rc |= gnb_make_tmap_bf16(ctx, &st->m_k, st->k, 3, d, s, box); }
    { const uint32_t bv[3] = {64, 64, 1};
      const uint64_t d[3] = {(uint64_t)kc, 64, slots * LG_HEADS}, s[2] = {(uint64_t)kc * 2, (uint64_t)kc * 64 * 2};
      rc |= gnb_make_tmap_bf16(ctx, &st->m_vt, st->vt, 3, d, s, bv); }
    if (rc) { gnb_lightglue_free(ctx); return GNB_E_CUDA; }
    GNB_CUDA(ctx, cudaFuncSetAttribute(lg_linear_persist_kernel<LG_QKV, 256, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, LgpCfg<256, 256>::SMEM));
    GNB_CUDA(ctx, cudaFuncSetAttribute(lg_linear_persist_kernel<LG_OUT, 256, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, LgpCfg<256, 256>::SMEM));
    GNB_CUDA(ctx, cudaFuncSetAttribute(lg_linear_kernel<LG_FC1, 512, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, LglCfg<512>::SMEM));
    GNB_CUDA(ctx, cudaFuncSetAttribute(lg_linear_persist_kernel<LG_FC2, 512, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, LgpCfg<512, 128>::SMEM));
    GNB_CUDA(ctx, cudaFuncSetAttribute(lg_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LGA_SMEM));
    return GNB_OK;
}

// Run the layers in place on ctx->desc_f32 for `pairs` (slot_a0 + p, slot_b0 + p); image sizes are (h, w) per side.
int gnb_lightglue_forward(gnb_ctx* ctx, int pairs, int slot_a0, int slot_b0, float ha, float wa, float hb, float wb) {
    LgState* st = lg_state(ctx);
    if (!st || pairs < 1) return GNB_OK;
    const int kc = ctx->cfg.max_keypoints;
    int* err = gnb_tc_err_dev(ctx);
    LgPairs pr{slot_a0, slot_b0, pairs};
    LgPos pos;
    memcpy(pos.w, st->pos_w, sizeof(pos.w));
    const int rt = ceil_div(kc, 128), z = 2 * pairs;
    // persistent linear kernels: CTAs per weight slice (the slices of one launch share the SMs)
    const int tiles = z * rt, sms = ctx->sm_count;
    const int cta1 = tiles < sms ? tiles : sms, cta2 = tiles < sms / 2 ? tiles : sms / 2, cta3 = tiles < sms / 3 ? tiles : sms / 3;
    GNB_KERNEL(ctx, "lg_init_kernel", lg_init_kernel<<<dim3(ceil_div(kc, 8), z), 256, 0, ctx->stream>>>(
        ctx->desc_f32, ctx->kp_xy, ctx->kp_count, pr, kc, pos, ha, wa, hb, wb, st->xo, st->cs));
    for (int l = 0; l < st->n_layers; ++l) {
        for (int blk = 0; blk < 2; ++blk) {
            const LgBlockW& w = st->blocks[2 * l + blk];
            GNB_KERNEL(ctx, "lg_linear<qkv>", lg_linear_persist_kernel<LG_QKV, 256, 256><<<dim3(cta3, 3), LGP_THREADS, LgpCfg<256, 256>::SMEM, ctx->stream>>>(
                st->m_x, w.m_qkv, ctx->kp_count, pr, kc, rt, w.bqkv, blk == 0, st->cs, st->q, st->k, st->vt, nullptr, err));
            GNB_KERNEL(ctx, blk == 0 ? "lg_attn<self>" : "lg_attn<cross>", lg_attn_kernel<<<dim3(rt, LG_HEADS, z), LGA_THREADS, LGA_SMEM, ctx->stream>>>(
                st->m_q, st->m_k, st->m_vt, ctx->kp_count, pr, kc, blk, st->att, err));
            GNB_KERNEL(ctx, "lg_linear<out>", lg_linear_persist_kernel<LG_OUT, 256, 256><<<dim3(cta1, 1), LGP_THREADS, LgpCfg<256, 256>::SMEM, ctx->stream>>>(
                st->m_att, w.m_o, ctx->kp_count, pr, kc, rt, w.bo, 0, nullptr, st->xo, nullptr, nullptr, nullptr, err));
            GNB_KERNEL(ctx, "lg_linear<fc1>", lg_linear_kernel<LG_FC1, 512, 512><<<dim3(rt, 1, z), LGL_THREADS(512), LglCfg<512>::SMEM, ctx->stream>>>(
                st->m_xo, w.m_w1, ctx->kp_count, pr, kc, w.b1, w.lng, w.lnb, 0, nullptr, st->h, nullptr, nullptr, nullptr, err));
            GNB_KERNEL(ctx, "lg_linear<fc2>", lg_linear_persist_kernel<LG_FC2, 512, 128><<<dim3(cta2, 2), LGP_THREADS, LgpCfg<512, 128>::SMEM, ctx->stream>>>(
                st->m_h, w.m_w2, ctx->kp_count, pr, kc, rt, w.b2, 0, nullptr, st->xo, nullptr, nullptr, ctx->desc_f32, err));
        }
    }
    return GNB_OK;
}
