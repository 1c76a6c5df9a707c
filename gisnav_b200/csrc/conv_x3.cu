// conv_x3.cu — the CUDA-core pieces of the fp32-faithful mode (cfg.precision = 1; contract in
// oracle/superpoint_ref.py, quantize = "x3"; reference tensors are fp32 end to end,
// ros/gisnav/gisnav/core/pose_node.py:254-287).
//
// In this mode every activation travels as two bf16 terms, v = hi + lo (16 significant bits), stored NHWC as
// channel blocks [hi: C | lo: C]; the 3x3 layers with Cin >= 64 run on tcgen05 with three MMAs per product
// (conv_tc.cu, X3 kernels).  What is left runs here in plain fp32 on the CUDA cores because it is < 1 % of the FLOPs:
//   conv1a        (Cin = 1: nine FMAs per output, write-bound)          u8 image -> [hi|lo] x 64 channels
//   the matcher head's projection and the dense descriptor map of the parity hook as a generic SIMT GEMM (K = 256)
// (the two 1x1 heads of the product path run on tcgen05 with the same split: heads_tc.cu)
#include "common.cuh"

#include <math.h>

// ------------------------------------------------------------------------------------------------
// conv1a, fp32: same tiling as conv1a_kernel (conv.cu).  Output pixel = 128 bf16: hi[64] then lo[64].
// The kernel is issue-bound (ncu: issue slots 73 % busy, DRAM 60 %; under the power cap of the fp32-faithful step the SM clock
// drops to ~1.6 GHz and an issue-bound kernel slows with it), so the arithmetic uses Blackwell's packed fp32 pairs — FFMA2 /
// FADD2 (`fma.rn.f32x2`, `add.rn.f32x2`): two IEEE fp32 operations per issue slot, bit-identical to the scalar form; the
// pixel value is a broadcast operand (`FFMA2 R, Rv.F32, Rw.F32x2.HI_LO, Racc`).  Measured per 128 frames: 7.79 -> 7.31 ms.
// A variant with eight channels per lane (16-byte stores, 124 registers, two CTAs per SM) was slower: 8.82 ms.
#define X1_TH 8
#define X1_TW 32
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

__global__ void __launch_bounds__(256, 4) conv1a_x3_kernel(const uint8_t* __restrict__ img, const float* __restrict__ wt,
                                                           const float* __restrict__ bias, int h, int w, bf16* __restrict__ out) {
    __shared__ float patch[X1_TH + 2][X1_TW + 2 + 2];
    const int b = blockIdx.z, y0 = blockIdx.y * X1_TH, x0 = blockIdx.x * X1_TW;
    const uint8_t* im = img + (size_t)b * h * w;
    for (int i = threadIdx.x; i < (X1_TH + 2) * (X1_TW + 2); i += 256) {
        const int ly = i / (X1_TW + 2), lx = i % (X1_TW + 2);
        const int y = y0 + ly - 1, x = x0 + lx - 1;
        float f = 0.f;
        if (y >= 0 && y < h && x >= 0 && x < w) f = __fdiv_rn((float)im[(size_t)y * w + x], 255.0f);
        patch[ly][lx] = f;
    }
    // lane l owns channels [4 (l % 16), +4) of pixel (l / 16) of a 2-pixel group: its 36 weights live in registers as 18
    // pairs (four CTAs per SM), and the 16 lanes of a pixel write its 128-byte hi block and its 128-byte lo block with one
    // 8-byte store each
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = (lane & 15) * 4, sub = lane >> 4;
    f32x2 wr[9][2], br[2];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 2; ++j) wr[t][j] = f2_pack(wt[t * 64 + c0 + 2 * j], wt[t * 64 + c0 + 2 * j + 1]);   // w_f32 [tap][cout_pad = 64][cin = 1]
#pragma unroll
    for (int j = 0; j < 2; ++j) br[j] = f2_pack(bias[c0 + 2 * j], bias[c0 + 2 * j + 1]);
    __syncthreads();
    const int ly = warp, y = y0 + ly;
    if (y >= h) return;
#pragma unroll 4
    for (int it = 0; it < X1_TW / 2; ++it) {
        const int lx = it * 2 + sub, x = x0 + lx;
        float v[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) v[t] = patch[ly + t / 3][lx + t % 3];
        uint32_t hi[2], lo[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            f32x2 acc = f2_pack(0.f, 0.f);
#pragma unroll
            for (int t = 0; t < 9; ++t) acc = f2_fma(f2_pack(v[t], v[t]), wr[t][j], acc);   // same chain per channel as the scalar form
            float a0, a1;
            f2_unpack(f2_add(acc, br[j]), a0, a1);
            a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f);
            const __nv_bfloat162 hh = __floats2bfloat162_rn(a0, a1);
            float r0, r1;
            f2_unpack(f2_sub(f2_pack(a0, a1), f2_pack(__low2float(hh), __high2float(hh))), r0, r1);
            const __nv_bfloat162 ll = __floats2bfloat162_rn(r0, r1);
            hi[j] = *reinterpret_cast<const uint32_t*>(&hh);
            lo[j] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        if (x < w) {
            bf16* o = out + ((size_t)b * h * w + (size_t)y * w + x) * 128 + c0;
            *reinterpret_cast<uint2*>(o) = make_uint2(hi[0], hi[1]);
            *reinterpret_cast<uint2*>(o + 64) = make_uint2(lo[0], lo[1]);
        }
    }
}

int gnb_conv1a_x3(gnb_ctx* ctx, const uint8_t* img, int n, int h, int w, bf16* out) {
    dim3 grid(ceil_div(w, X1_TW), ceil_div(h, X1_TH), n);
    GNB_KERNEL(ctx, "conv1a_x3_kernel", conv1a_x3_kernel<<<grid, 256, 0, ctx->stream>>>(img, ctx->layers[L1A].w_f32, ctx->layers[L1A].bias, h, w, out));
    return GNB_OK;
}

// ------------------------------------------------------------------------------------------------
// Generic fp32 SIMT GEMM with K = 256:  C[m][n] = scale * (bias[n] + sum_k A[m][k] W[n][k]).
//   AMODE 0: A rows are fp32 [256]            (matcher head projection of the descriptors)
//   AMODE 1: A rows are split activations: 256 bf16 hi then 256 bf16 lo; a = float(hi) + float(lo)
// row_idx (optional): source row of output row m (-1 = all-zero row, output row = scale * bias); NULL = identity.
// CTA = 64 x 64 output tile, 256 threads, 4 x 4 micro-tile per thread, K in steps of 32 through shared memory.
template <int AMODE>
__global__ void __launch_bounds__(256) gemm256_f32_kernel(const void* __restrict__ a, const int* __restrict__ row_idx, int m_rows,
                                                          const float* __restrict__ wgt, const float* __restrict__ bias, int n_cols,
                                                          float scale, float* __restrict__ c, int ldc) {
    __shared__ float As[32][64 + 4];   // [k][m]
    __shared__ float Ws[32][64 + 4];   // [k][n]
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    f32x2 acc[4][2];   // packed pairs along n: FFMA2 halves the issue slots of the inner loop, same fp32 chain per output
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j] = f2_pack(0.f, 0.f);
    // loader mapping: thread -> (row r = tid / 4, 8 consecutive k = (tid % 4) * 8)
    const int lr = tid >> 2, lk = (tid & 3) * 8;
    int src = -1;
    if (m0 + lr < m_rows) src = row_idx ? row_idx[m0 + lr] : m0 + lr;
    const bool wrow_ok = n0 + lr < n_cols;
    for (int k0 = 0; k0 < 256; k0 += 32) {
        float av[8], wv[8];
        if (src >= 0) {
            if (AMODE == 0) {
                const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(a) + (size_t)src * 256 + k0 + lk);
                const float4 u = __ldg(p), v = __ldg(p + 1);
                av[0] = u.x; av[1] = u.y; av[2] = u.z; av[3] = u.w; av[4] = v.x; av[5] = v.y; av[6] = v.z; av[7] = v.w;
            } else {
                const bf16* row = static_cast<const bf16*>(a) + (size_t)src * 512 + k0 + lk;
                const uint4 hi = __ldg(reinterpret_cast<const uint4*>(row)), lo = __ldg(reinterpret_cast<const uint4*>(row + 256));
                const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    av[2 * j] = __fadd_rn(__uint_as_float(hw[j] << 16), __uint_as_float(lw[j] << 16));
                    av[2 * j + 1] = __fadd_rn(__uint_as_float(hw[j] & 0xffff0000u), __uint_as_float(lw[j] & 0xffff0000u));
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) av[j] = 0.f;
        }
        if (wrow_ok) {
            const float4* p = reinterpret_cast<const float4*>(wgt + (size_t)(n0 + lr) * 256 + k0 + lk);
            const float4 u = __ldg(p), v = __ldg(p + 1);
            wv[0] = u.x; wv[1] = u.y; wv[2] = u.z; wv[3] = u.w; wv[4] = v.x; wv[5] = v.y; wv[6] = v.z; wv[7] = v.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) wv[j] = 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) { As[lk + j][lr] = av[j]; Ws[lk + j][lr] = wv[j]; }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const float4 x = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 y = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
            const float xa[4] = {x.x, x.y, x.z, x.w};
            const f32x2 yb[2] = {f2_pack(y.x, y.y), f2_pack(y.z, y.w)};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) acc[i][j] = f2_fma(f2_pack(xa[i], xa[i]), yb[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= m_rows) continue;
        float r[4];
        f2_unpack(acc[i][0], r[0], r[1]);
        f2_unpack(acc[i][1], r[2], r[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < n_cols) c[(size_t)m * ldc + n] = __fmul_rn(__fadd_rn(r[j], bias[n]), scale);
        }
    }
}

int gnb_gemm256_f32(gnb_ctx* ctx, int amode, const void* a, const int* row_idx, int m_rows, const float* wgt, const float* bias, int n_cols,
                    float scale, float* c, int ldc, const char* name) {
    if (m_rows <= 0) return GNB_OK;
    dim3 grid(ceil_div(m_rows, 64), ceil_div(n_cols, 64));
    if (amode == 0)
        GNB_KERNEL(ctx, name, gemm256_f32_kernel<0><<<grid, 256, 0, ctx->stream>>>(a, row_idx, m_rows, wgt, bias, n_cols, scale, c, ldc));
    else
        GNB_KERNEL(ctx, name, gemm256_f32_kernel<1><<<grid, 256, 0, ctx->stream>>>(a, row_idx, m_rows, wgt, bias, n_cols, scale, c, ldc));
    return GNB_OK;
}

// ------------------------------------------------------------------------------------------------
// parity hook: split activation [pixels][hi: c | lo: c] -> f32 [pixels][c]
__global__ void split_to_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out, size_t pixels, int c) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pixels * c) return;
    const size_t p = i / c;
    const int ch = (int)(i - p * c);
    out[i] = __fadd_rn(__bfloat162float(in[p * 2 * c + ch]), __bfloat162float(in[p * 2 * c + c + ch]));
}

int gnb_split_to_f32(gnb_ctx* ctx, const bf16* in, float* out, size_t pixels, int c) {
    const size_t n = pixels * c;
    GNB_KERNEL(ctx, "split_to_f32_kernel", split_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(in, out, pixels, c));
    return GNB_OK;
}

// matchability logit in fp32: one warp per keypoint
__global__ void __launch_bounds__(256) mlogit_f32_kernel(const float* __restrict__ desc, const int* __restrict__ kp_count, int slot0, int k_cap,
                                                         const float* __restrict__ mw, float mb, float* __restrict__ mlogit) {
    const int slot = slot0 + blockIdx.y;
    const int kp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (kp >= kp_count[slot]) return;
    const float* d = desc + ((size_t)slot * k_cap + kp) * 256;
    float z = 0.f;
    for (int i = lane; i < 256; i += 32) z = fmaf(d[i], mw[i], z);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) z += __shfl_xor_sync(0xffffffffu, z, s);
    z += mb;
    if (lane == 0) mlogit[(size_t)slot * k_cap + kp] = fminf(z, 0.f) - log1pf(expf(-fabsf(z)));
}

// fp32 projected descriptors -> two bf16 terms per value, [row][hi: 256 | lo: 256] (operands of match_pair_tc<.., X3>)
__global__ void split_rows_kernel(const float* __restrict__ in, bf16* __restrict__ out, size_t rows) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * 256) return;
    const size_t r = i >> 8;
    const int c = (int)(i & 255);
    const float v = in[i];
    const bf16 hi = __float2bfloat16_rn(v);
    out[r * 512 + c] = hi;
    out[r * 512 + 256 + c] = __float2bfloat16_rn(__fsub_rn(v, __bfloat162float(hi)));
}

int gnb_project_f32(gnb_ctx* ctx, int slot0, int n_slots) {
    const int k = ctx->cfg.max_keypoints;
    // rows beyond a slot's keypoint count are projected too (finite garbage-free: the buffers are zero-initialised
    // and only rows < count are ever read by the matcher)
    int rc = gnb_gemm256_f32(ctx, 0, ctx->desc_f32 + (size_t)slot0 * k * 256, nullptr, n_slots * k, ctx->match_w_f32, ctx->match_b, 256, 0.25f,
                             ctx->mproj_f32 + (size_t)slot0 * k * 256, 256, "project_f32");
    if (rc) return rc;
    if (ctx->mproj_x3) {
        const size_t rows = (size_t)n_slots * k;
        GNB_KERNEL(ctx, "split_rows_kernel", split_rows_kernel<<<(unsigned)((rows * 256 + 255) / 256), 256, 0, ctx->stream>>>(
            ctx->mproj_f32 + (size_t)slot0 * k * 256, ctx->mproj_x3 + (size_t)slot0 * k * 512, rows));
    }
    dim3 grid(ceil_div(k * 32, 256), n_slots);
    GNB_KERNEL(ctx, "mlogit_f32_kernel", mlogit_f32_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->desc_f32, ctx->kp_count, slot0, k, ctx->match_mw_f32,
                                                                                   ctx->match_mb, ctx->mlogit));
    return GNB_OK;
}
