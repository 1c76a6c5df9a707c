// conv.cu — K1: the SuperPoint-style dense stack (replaces cv2.SIFT.detectAndCompute,
// ros/gisnav/gisnav/core/pose_node.py:230; architecture per SURVEY.md §8(c)).
//
// Numerics contract (oracle/superpoint_ref.py): conv operands (activations, weights) are bf16,
// accumulation fp32, bias fp32, ReLU, activations stored back as bf16 NHWC.  Heads stay fp32.
//
// This file holds weight repacking, the Cin=1 first layer, the detector/descriptor head epilogues
// and a SIMT validation implementation of the 3x3/1x1 convs (cfg.conv_impl = 1).  The product
// path for the Cin>=64 layers is the tcgen05 implicit GEMM in conv_tc.cu.
#include "common.cuh"

#include <stdlib.h>

struct CUtensorMap_st;
const CUtensorMap_st* gnb_conv_tc_wmap(gnb_ctx* ctx, int lid);
const CUtensorMap_st* gnb_conv_tc_wmap_x3(gnb_ctx* ctx, int lid);
int gnb_score_head_tc(gnb_ctx* ctx, const CUtensorMap_st* tmap_w, const bf16* apa, const float* bias, int n, int hc, int wc, float* score);
int gnb_conv1_fused_tc(gnb_ctx* ctx, const uint8_t* img, int n, int h, int w, bf16* out_p1);
int gnb_desc_head_tc(gnb_ctx* ctx, const CUtensorMap_st* tmap_w, const bf16* ada, const float* bias, int n, int h, int w, int slot0);
// conv_x3.cu (fp32-faithful mode)
int gnb_conv1a_x3(gnb_ctx* ctx, const uint8_t* img, int n, int h, int w, bf16* out);
int gnb_gemm256_f32(gnb_ctx* ctx, int amode, const void* a, const int* row_idx, int m_rows, const float* wgt, const float* bias, int n_cols,
                    float scale, float* c, int ldc, const char* name);
int gnb_desc_head_x3_tc(gnb_ctx* ctx, const CUtensorMap_st* tmap_w, const bf16* ada, const float* bias, int n, int h, int w, int slot0);

struct LayerSpec { const char* name; int cin, cout, ks; };
static const LayerSpec kSpecs[GNB_NUM_LAYERS] = {
    {"conv1a", 1, 64, 3},   {"conv1b", 64, 64, 3},   {"conv2a", 64, 64, 3},   {"conv2b", 64, 64, 3},
    {"conv3a", 64, 128, 3}, {"conv3b", 128, 128, 3}, {"conv4a", 128, 128, 3}, {"conv4b", 128, 128, 3},
    {"convPa", 128, 256, 3}, {"convPb", 256, 65, 1}, {"convDa", 128, 256, 3}, {"convDb", 256, 256, 1},
};

// Weight repack ON THE DEVICE: blob tensor [cout][cin][ks][ks] f32 (PyTorch order) ->
//   w     bf16 [tap][cout_pad][cin]            (K-major B operand per tap)
//   w_x3  bf16 [tap][cout_pad][hi: cin | lo: cin]   (fp32-faithful mode: w = hi + lo)
//   w_f32 f32  [tap][cout_pad][cin]            (fp32-faithful mode, CUDA-core layers)
// __float2bfloat16_rn is round-to-nearest-even, identical to torch's .to(bfloat16) the oracle uses.
__global__ void repack_conv_kernel(const float* __restrict__ w, int cout, int cin, int taps, int cout_pad, bf16* __restrict__ out,
                                   bf16* __restrict__ out_x3, float* __restrict__ out_f32) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // over [tap][cout_pad][cin]
    const size_t total = (size_t)taps * cout_pad * cin;
    if (i >= total) return;
    const int ci = (int)(i % cin), co = (int)((i / cin) % cout_pad), t = (int)(i / ((size_t)cin * cout_pad));
    const float v = co < cout ? w[((size_t)co * cin + ci) * taps + t] : 0.f;
    const bf16 hi = __float2bfloat16_rn(v);
    out[i] = hi;
    if (out_x3) {
        const size_t row = ((size_t)t * cout_pad + co) * 2 * cin;
        out_x3[row + ci] = hi;
        out_x3[row + cin + ci] = __float2bfloat16_rn(__fsub_rn(v, __bfloat162float(hi)));
    }
    if (out_f32) out_f32[i] = v;
}

__global__ void pad_bias_kernel(const float* __restrict__ b, int cout, int cout_pad, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cout_pad) out[i] = i < cout ? b[i] : 0.f;
}

int gnb_conv_init(gnb_ctx* ctx, const float* blob) {
    size_t off = 0;
    for (int l = 0; l < GNB_NUM_LAYERS; ++l) {
        const LayerSpec& s = kSpecs[l];
        ConvLayer& L = ctx->layers[l];
        L.cin = s.cin; L.cout = s.cout; L.ks = s.ks;
        L.cout_pad = (s.cout + 31) / 32 * 32;  // tcgen05 epilogue reads 32 columns at a time
        const int taps = s.ks * s.ks;
        const float* w = blob + off;                       // [cout][cin][ks][ks] (device)
        const float* b = w + (size_t)s.cout * s.cin * taps;  // [cout]
        off += (size_t)s.cout * s.cin * taps + s.cout;
        const size_t n = (size_t)taps * L.cout_pad * s.cin;
        GNB_CUDA(ctx, cudaMalloc(&L.w, n * sizeof(bf16)));
        GNB_CUDA(ctx, cudaMalloc(&L.bias, L.cout_pad * sizeof(float)));
        if (ctx->cfg.precision == 1) {
            GNB_CUDA(ctx, cudaMalloc(&L.w_x3, 2 * n * sizeof(bf16)));
            GNB_CUDA(ctx, cudaMalloc(&L.w_f32, n * sizeof(float)));
        }
        GNB_KERNEL(ctx, "repack_conv_kernel", repack_conv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(w, s.cout, s.cin, taps, L.cout_pad, L.w, L.w_x3, L.w_f32));
        GNB_KERNEL(ctx, "pad_bias_kernel", pad_bias_kernel<<<ceil_div(L.cout_pad, 128), 128, 0, ctx->stream>>>(b, s.cout, L.cout_pad, L.bias));
    }
    return GNB_OK;
}

void gnb_conv_free(gnb_ctx* ctx) {
    for (int l = 0; l < GNB_NUM_LAYERS; ++l) {
        if (ctx->layers[l].w) cudaFree(ctx->layers[l].w);
        if (ctx->layers[l].bias) cudaFree(ctx->layers[l].bias);
        if (ctx->layers[l].w_x3) cudaFree(ctx->layers[l].w_x3);
        if (ctx->layers[l].w_f32) cudaFree(ctx->layers[l].w_f32);
        ctx->layers[l].w = nullptr; ctx->layers[l].bias = nullptr; ctx->layers[l].w_x3 = nullptr; ctx->layers[l].w_f32 = nullptr;
    }
}

// ------------------------------------------------------------------------------------------------
// conv1a: u8 image -> 64 channels, 3x3, ReLU, bf16 NHWC.  Cin = 1, so this is 9 FMAs per output and
// the kernel is HBM-write-bound (1 B in, 128 B out per pixel).  CTA = 8 x 32 pixel tile; lane l of a
// warp owns channels [8 (l%8), +8) of pixel (l/8) of a 4-pixel group, so every warp-wide 16-byte
// store covers 512 contiguous bytes.  The 72 weights a thread needs live in registers.
#define C1_TH 8
#define C1_TW 32
__global__ void __launch_bounds__(256) conv1a_kernel(const uint8_t* __restrict__ img, const bf16* __restrict__ wt,
                                                     const float* __restrict__ bias, int h, int w, bf16* __restrict__ out) {
    __shared__ float patch[C1_TH + 2][C1_TW + 2 + 2];
    const int b = blockIdx.z, y0 = blockIdx.y * C1_TH, x0 = blockIdx.x * C1_TW;
    const uint8_t* im = img + (size_t)b * h * w;
    for (int i = threadIdx.x; i < (C1_TH + 2) * (C1_TW + 2); i += 256) {
        const int ly = i / (C1_TW + 2), lx = i % (C1_TW + 2);
        const int y = y0 + ly - 1, x = x0 + lx - 1;
        float f = 0.f;
        if (y >= 0 && y < h && x >= 0 && x < w) f = (float)im[(size_t)y * w + x] / 255.0f;
        patch[ly][lx] = __bfloat162float(__float2bfloat16_rn(f));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = (lane & 7) * 8, sub = lane >> 3;
    float wr[9][8], br[8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) wr[t][j] = __bfloat162float(wt[t * 64 + c0 + j]);
#pragma unroll
    for (int j = 0; j < 8; ++j) br[j] = bias[c0 + j];
    __syncthreads();
    // warp `warp` owns row y0 + warp; 8 iterations of 4 pixels
    const int ly = warp, y = y0 + ly;
    if (y >= h) return;
#pragma unroll 2
    for (int it = 0; it < C1_TW / 4; ++it) {
        const int lx = it * 4 + sub, x = x0 + lx;
        float v[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) v[t] = patch[ly + t / 3][lx + t % 3];
        __align__(16) __nv_bfloat162 r[4];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int t = 0; t < 9; ++t) { a0 = fmaf(v[t], wr[t][j], a0); a1 = fmaf(v[t], wr[t][j + 1], a1); }
            a0 += br[j]; a1 += br[j + 1];
            r[j / 2] = __floats2bfloat162_rn(fmaxf(a0, 0.f), fmaxf(a1, 0.f));
        }
        if (x < w) *reinterpret_cast<uint4*>(out + ((size_t)b * h * w + (size_t)y * w + x) * 64 + c0) = *reinterpret_cast<const uint4*>(r);
    }
}

// ------------------------------------------------------------------------------------------------
// SIMT validation conv: 8x16 pixel tile per CTA, one pixel per thread, 16 output channels at a time.
template <int CIN, int KS>
__global__ void __launch_bounds__(128) conv_simt_kernel(const bf16* __restrict__ in, const bf16* __restrict__ wt,
                                                        const float* __restrict__ bias, int h, int w, int cout,
                                                        int cout_pad, bf16* __restrict__ out_bf,
                                                        float* __restrict__ out_f, int relu, int pool) {
    constexpr int R = KS / 2, TH = 8, TW = 16, HW = TW + 2 * R, HH = TH + 2 * R;
    constexpr int CP = CIN + 2;  // padded pixel pitch (bf16 elements) -> conflict-free column reads
    extern __shared__ __align__(16) uint8_t smem_raw[];
    bf16* tile = reinterpret_cast<bf16*>(smem_raw);
    float* wsm = reinterpret_cast<float*>(smem_raw + ((HH * HW * CP * 2 + 15) / 16) * 16);
    const int tid = threadIdx.x, b = blockIdx.z, y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
    const bf16* inb = in + (size_t)b * h * w * CIN;
    for (int i = tid; i < HH * HW * (CIN / 2); i += 128) {
        int p = i / (CIN / 2), v = i % (CIN / 2);
        int yy = y0 - R + p / HW, xx = x0 - R + p % HW;
        uint32_t val = 0;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w)
            val = *reinterpret_cast<const uint32_t*>(inb + ((size_t)yy * w + xx) * CIN + v * 2);
        *reinterpret_cast<uint32_t*>(tile + p * CP + v * 2) = val;
    }
    const int ty = tid / TW, tx = tid % TW;
    const int y = y0 + ty, x = x0 + tx;
    for (int c0 = 0; c0 < cout_pad; c0 += 16) {
        __syncthreads();
        for (int i = tid; i < KS * KS * 16 * CIN; i += 128) {
            int ci = i % CIN, j = (i / CIN) % 16, t = i / (CIN * 16);
            wsm[(t * CIN + ci) * 16 + j] = __bfloat162float(wt[((size_t)t * cout_pad + c0 + j) * CIN + ci]);
        }
        __syncthreads();
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
        for (int t = 0; t < KS * KS; ++t) {
            const bf16* a = tile + ((ty + t / KS) * HW + tx + t % KS) * CP;
            const float* wrow = wsm + (size_t)t * CIN * 16;
#pragma unroll 4
            for (int ci = 0; ci < CIN; ci += 2) {
                __nv_bfloat162 av = *reinterpret_cast<const __nv_bfloat162*>(a + ci);
                float a0 = __low2float(av), a1 = __high2float(av);
                const float4* w0 = reinterpret_cast<const float4*>(wrow + ci * 16);
                const float4* w1 = w0 + 4;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4 wa = w0[q], wb = w1[q];
                    acc[4 * q + 0] = fmaf(a0, wa.x, acc[4 * q + 0]);
                    acc[4 * q + 1] = fmaf(a0, wa.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(a0, wa.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(a0, wa.w, acc[4 * q + 3]);
                    acc[4 * q + 0] = fmaf(a1, wb.x, acc[4 * q + 0]);
                    acc[4 * q + 1] = fmaf(a1, wb.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(a1, wb.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(a1, wb.w, acc[4 * q + 3]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float v = acc[j] + bias[c0 + j];
            if (relu) v = fmaxf(v, 0.f);
            if (pool) {
                v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
                v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
            }
            acc[j] = v;
        }
        if (pool) {
            if (((tx | ty) & 1) == 0 && y < h && x < w) {
                const int ho = h / 2, wo = w / 2;
                bf16* o = out_bf + (((size_t)b * ho + y / 2) * wo + x / 2) * cout + c0;
                for (int j = 0; j < 16 && c0 + j < cout; ++j) o[j] = __float2bfloat16_rn(acc[j]);
            }
        } else if (y < h && x < w) {
            size_t pix = ((size_t)b * h + y) * w + x;
            if (out_bf) {
                bf16* o = out_bf + pix * cout + c0;
                for (int j = 0; j < 16 && c0 + j < cout; ++j) o[j] = __float2bfloat16_rn(acc[j]);
            } else {
                float* o = out_f + pix * cout + c0;
                for (int j = 0; j < 16 && c0 + j < cout; ++j) o[j] = acc[j];
            }
        }
    }
}

template <int CIN, int KS>
static int launch_conv_simt(gnb_ctx* ctx, const ConvLayer& L, const bf16* in, int n, int h, int w, bf16* out_bf,
                            float* out_f, int relu, int pool) {
    constexpr int R = KS / 2;
    size_t smem = (((8 + 2 * R) * (16 + 2 * R) * (CIN + 2) * 2 + 15) / 16) * 16 + (size_t)KS * KS * 16 * CIN * 4;
    GNB_CUDA(ctx, gnb_func_smem(ctx, conv_simt_kernel<CIN, KS>, (int)smem));
    dim3 grid(ceil_div(w, 16), ceil_div(h, 8), n);
    GNB_KERNEL(ctx, "conv_simt_kernel", conv_simt_kernel<CIN, KS><<<grid, 128, smem, ctx->stream>>>(in, L.w, L.bias, h, w, L.cout, L.cout_pad, out_bf, out_f, relu, pool));
    return GNB_OK;
}

static int conv_layer(gnb_ctx* ctx, int lid, const bf16* in, int n, int h, int w, bf16* out_bf, float* out_f,
                      int relu, int pool) {
    const ConvLayer& L = ctx->layers[lid];
    if (ctx->cfg.conv_impl == 0 || ctx->cfg.precision == 1) {
        int rc = gnb_conv_tc_layer(ctx, L, in, n, h, w, out_bf, out_f, relu, pool);
        if (rc != GNB_E_INVALID) return rc;
        GNB_SET_ERR(ctx, "tcgen05 conv does not support layer %d (cin %d cout %d ks %d)", lid, L.cin, L.cout, L.ks);
        return rc;
    }
    if (L.cin == 64 && L.ks == 3) return launch_conv_simt<64, 3>(ctx, L, in, n, h, w, out_bf, out_f, relu, pool);
    if (L.cin == 128 && L.ks == 3) return launch_conv_simt<128, 3>(ctx, L, in, n, h, w, out_bf, out_f, relu, pool);
    if (L.cin == 256 && L.ks == 1) return launch_conv_simt<256, 1>(ctx, L, in, n, h, w, out_bf, out_f, relu, pool);
    GNB_SET_ERR(ctx, "unsupported conv layer shape");
    return GNB_E_INVALID;
}

// ------------------------------------------------------------------------------------------------
// detector head epilogue: softmax over 65 logits per cell, drop dustbin, 8x8 depth-to-space.
// HBM-bound: 260 B in, 256 B out per cell.  One thread per cell.
__global__ void __launch_bounds__(128) softmax_d2s_kernel(const float* __restrict__ semi, int n_cells_total, int hc,
                                                          int wc, float* __restrict__ score) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= n_cells_total) return;
    const float* s = semi + (size_t)cell * 65;
    float v[65];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 65; ++i) { v[i] = s[i]; m = fmaxf(m, v[i]); }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 65; ++i) { v[i] = expf(v[i] - m); sum += v[i]; }
    const float inv = 1.0f / sum;
    const int b = cell / (hc * wc), rem = cell % (hc * wc), cy = rem / wc, cx = rem % wc;
    const int w = wc * 8;
    float* o = score + ((size_t)b * hc * 8 + (size_t)cy * 8) * w + (size_t)cx * 8;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        float4 lo = make_float4(v[r * 8 + 0] * inv, v[r * 8 + 1] * inv, v[r * 8 + 2] * inv, v[r * 8 + 3] * inv);
        float4 hi = make_float4(v[r * 8 + 4] * inv, v[r * 8 + 5] * inv, v[r * 8 + 6] * inv, v[r * 8 + 7] * inv);
        *reinterpret_cast<float4*>(o + (size_t)r * w) = lo;
        *reinterpret_cast<float4*>(o + (size_t)r * w + 4) = hi;
    }
}

// descriptor head epilogue: L2-normalise 256 channels per cell in place. One warp per cell.
__global__ void __launch_bounds__(256) l2norm256_kernel(float* __restrict__ d, int n_cells_total) {
    const int cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (cell >= n_cells_total) return;
    float4* p = reinterpret_cast<float4*>(d + (size_t)cell * 256);
    float4 a = p[lane], b = p[lane + 32];
    float ss = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv; b.x *= inv; b.y *= inv; b.z *= inv; b.w *= inv;
    p[lane] = a; p[lane + 32] = b;
}

// ------------------------------------------------------------------------------------------------
int gnb_conv_forward(gnb_ctx* ctx, int n, int h, int w, int dense_desc) {
    ConvWorkspace& cw = ctx->cw;
    if (n > cw.cap_images || (size_t)h * w > cw.cap_pixels || (h % 8) || (w % 8)) {
        GNB_SET_ERR(ctx, "conv_forward: %d images of %dx%d exceed the workspace or are not multiples of 8", n, h, w);
        return GNB_E_CAPACITY;
    }
    cw.n = n; cw.h = h; cw.w = w;
    int rc;
    if (ctx->cfg.precision == 1) {
        // fp32-faithful mode: conv1a fp32 (CUDA cores) -> split-bf16 tcgen05 stack -> fp32 heads
        // conv1a stays a separate fp32 kernel in this mode: fusing it into the conv1b kernel (producer warps computing
        // conv1a into the operand ring) measured SLOWER (36 vs 31 ms per 64 pairs): with hi + lo weights resident only
        // three 23 KB operand stages fit, so the producers cannot run a tile ahead of the MMA warp
        if ((rc = gnb_conv1a_x3(ctx, cw.img, n, h, w, cw.a1a))) return rc;
        if ((rc = conv_layer(ctx, L1B, cw.a1a, n, h, w, cw.p1, nullptr, 1, 1))) return rc;
        if ((rc = conv_layer(ctx, L2A, cw.p1, n, h / 2, w / 2, cw.a2a, nullptr, 1, 0))) return rc;
        if ((rc = conv_layer(ctx, L2B, cw.a2a, n, h / 2, w / 2, cw.p2, nullptr, 1, 1))) return rc;
        if ((rc = conv_layer(ctx, L3A, cw.p2, n, h / 4, w / 4, cw.a3a, nullptr, 1, 0))) return rc;
        if ((rc = conv_layer(ctx, L3B, cw.a3a, n, h / 4, w / 4, cw.p3, nullptr, 1, 1))) return rc;
        if ((rc = conv_layer(ctx, L4A, cw.p3, n, h / 8, w / 8, cw.a4a, nullptr, 1, 0))) return rc;
        if ((rc = conv_layer(ctx, L4B, cw.a4a, n, h / 8, w / 8, cw.a4b, nullptr, 1, 0))) return rc;
        if ((rc = conv_layer(ctx, LPA, cw.a4b, n, h / 8, w / 8, cw.apa, nullptr, 1, 0))) return rc;
        const int cells = n * (h / 8) * (w / 8);
        // detector head on tcgen05 with split operands, fused with softmax + depth-to-space (score_head_kernel<X3>)
        if ((rc = gnb_score_head_tc(ctx, gnb_conv_tc_wmap_x3(ctx, LPB), cw.apa, ctx->layers[LPB].bias, n, h / 8, w / 8, cw.score))) return rc;
        if ((rc = conv_layer(ctx, LDA, cw.a4b, n, h / 8, w / 8, cw.ada, nullptr, 1, 0))) return rc;
        if (dense_desc) {
            if ((rc = gnb_gemm256_f32(ctx, 1, cw.ada, nullptr, cells, ctx->layers[LDB].w_f32, ctx->layers[LDB].bias, 256, 1.0f, cw.dense, 256, "desc_dense_f32"))) return rc;
            GNB_KERNEL(ctx, "l2norm256_kernel", l2norm256_kernel<<<ceil_div(cells * 32, 256), 256, 0, ctx->stream>>>(cw.dense, cells));
        }
        return GNB_OK;
    }
    static const int no_fuse = getenv("GNB_NO_CONV1_FUSION") ? atoi(getenv("GNB_NO_CONV1_FUSION")) : 0;
    const bool fused1 = ctx->cfg.conv_impl == 0 && !no_fuse && (w % 16) == 0;  // TMA on the u8 image needs a 16-byte row pitch
    if (!fused1 || dense_desc) {
        // standalone conv1a: SIMT path, or the stage hooks that expose the conv1a activation
        dim3 grid(ceil_div(w, C1_TW), ceil_div(h, C1_TH), n);
        GNB_KERNEL(ctx, "conv1a_kernel", conv1a_kernel<<<grid, 256, 0, ctx->stream>>>(cw.img, ctx->layers[L1A].w, ctx->layers[L1A].bias, h, w, cw.a1a));
    }
    if (fused1) {
        if ((rc = gnb_conv1_fused_tc(ctx, cw.img, n, h, w, cw.p1))) return rc;
    } else {
        if ((rc = conv_layer(ctx, L1B, cw.a1a, n, h, w, cw.p1, nullptr, 1, 1))) return rc;
    }
    if ((rc = conv_layer(ctx, L2A, cw.p1, n, h / 2, w / 2, cw.a2a, nullptr, 1, 0))) return rc;
    if ((rc = conv_layer(ctx, L2B, cw.a2a, n, h / 2, w / 2, cw.p2, nullptr, 1, 1))) return rc;
    if ((rc = conv_layer(ctx, L3A, cw.p2, n, h / 4, w / 4, cw.a3a, nullptr, 1, 0))) return rc;
    if ((rc = conv_layer(ctx, L3B, cw.a3a, n, h / 4, w / 4, cw.p3, nullptr, 1, 1))) return rc;
    if ((rc = conv_layer(ctx, L4A, cw.p3, n, h / 8, w / 8, cw.a4a, nullptr, 1, 0))) return rc;
    if ((rc = conv_layer(ctx, L4B, cw.a4a, n, h / 8, w / 8, cw.a4b, nullptr, 1, 0))) return rc;
    if ((rc = conv_layer(ctx, LPA, cw.a4b, n, h / 8, w / 8, cw.apa, nullptr, 1, 0))) return rc;
    const int cells = n * (h / 8) * (w / 8);
    if (ctx->cfg.conv_impl == 0) {
        // detector head fused: 1x1 conv + softmax + depth-to-space straight into the score map
        if ((rc = gnb_score_head_tc(ctx, gnb_conv_tc_wmap(ctx, LPB), cw.apa, ctx->layers[LPB].bias, n, h / 8, w / 8, cw.score))) return rc;
    } else {
        if ((rc = conv_layer(ctx, LPB, cw.apa, n, h / 8, w / 8, nullptr, cw.semi, 0, 0))) return rc;
        GNB_KERNEL(ctx, "softmax_d2s_kernel", softmax_d2s_kernel<<<ceil_div(cells, 128), 128, 0, ctx->stream>>>(cw.semi, cells, h / 8, w / 8, cw.score));
    }
    if ((rc = conv_layer(ctx, LDA, cw.a4b, n, h / 8, w / 8, cw.ada, nullptr, 1, 0))) return rc;
    if (dense_desc || ctx->cfg.conv_impl != 0) {
        // dense 256-channel map (validation / stage hooks); the product path evaluates convDb on demand
        if ((rc = conv_layer(ctx, LDB, cw.ada, n, h / 8, w / 8, nullptr, cw.dense, 0, 0))) return rc;
        GNB_KERNEL(ctx, "l2norm256_kernel", l2norm256_kernel<<<ceil_div(cells * 32, 256), 256, 0, ctx->stream>>>(cw.dense, cells));
    }
    return GNB_OK;
}

int gnb_describe(gnb_ctx* ctx, int n, int h, int w, int slot0) {
    if (ctx->cfg.precision == 1)   // on demand, tcgen05 with split operands
        return gnb_desc_head_x3_tc(ctx, gnb_conv_tc_wmap_x3(ctx, LDB), ctx->cw.ada, ctx->layers[LDB].bias, n, h, w, slot0);
    if (ctx->cfg.conv_impl == 0)
        return gnb_desc_head_tc(ctx, gnb_conv_tc_wmap(ctx, LDB), ctx->cw.ada, ctx->layers[LDB].bias, n, h, w, slot0);
    return gnb_kp_sample(ctx, ctx->cw.dense, n, h, w, slot0);
}
