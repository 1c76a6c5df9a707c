// keypoints.cu — K2 (NMS + threshold + border + top-K) and K3 (descriptor sampling).
//
// K2 restates SuperPoint's simple_nms (two suppression refinements of a 9x9 max-pool), the score
// threshold, the four-sided border removal and a top-K whose order is DEFINED as ascending
// (-score, y*W+x) — oracle/nms_ref.py.  It is compare-only work, so the result is bit-identical
// to the oracle on the same score map.  Replaces the detector half of cv2.SIFT.detectAndCompute
// (ros/gisnav/gisnav/core/pose_node.py:230) and the keypoint list handling at pose_node.py:244-252.
//
// K3 restates SuperPoint's sample_descriptors (oracle/sample_ref.py).
#include "common.cuh"

#include <math.h>

#define NMS_TILE 32
#define NMS_HALO 20  // 5 pools x radius 4
#define NMS_REG (NMS_TILE + 2 * NMS_HALO)

// separable (2r+1)^2 max over a REG x REG smem array; window clipped to the array (== -inf pad)
__device__ __forceinline__ void pool_rows(const float* __restrict__ src, float* __restrict__ dst, int r) {
    for (int i = threadIdx.x; i < NMS_REG * NMS_REG; i += blockDim.x) {
        const int y = i / NMS_REG, x = i % NMS_REG;
        const int lo = max(x - r, 0), hi = min(x + r, NMS_REG - 1);
        float m = src[y * NMS_REG + lo];
        for (int xx = lo + 1; xx <= hi; ++xx) m = fmaxf(m, src[y * NMS_REG + xx]);
        dst[i] = m;
    }
}
__device__ __forceinline__ void pool_cols(const float* __restrict__ src, float* __restrict__ dst, int r) {
    for (int i = threadIdx.x; i < NMS_REG * NMS_REG; i += blockDim.x) {
        const int y = i / NMS_REG, x = i % NMS_REG;
        const int lo = max(y - r, 0), hi = min(y + r, NMS_REG - 1);
        float m = src[lo * NMS_REG + x];
        for (int yy = lo + 1; yy <= hi; ++yy) m = fmaxf(m, src[yy * NMS_REG + x]);
        dst[i] = m;
    }
}

__global__ void __launch_bounds__(256) nms_kernel(const float* __restrict__ score, int h, int w, int radius,
                                                  float threshold, int border, int slot0,
                                                  unsigned long long* __restrict__ cand_keys,
                                                  int* __restrict__ cand_count) {
    extern __shared__ float sm[];
    float* S = sm;                          // scores, -inf outside the image
    float* T = S + NMS_REG * NMS_REG;       // row-pass temp
    float* A = T + NMS_REG * NMS_REG;       // pooled result / suppressed scores
    float* F = A + NMS_REG * NMS_REG;       // mask as float
    uint8_t* M = reinterpret_cast<uint8_t*>(F + NMS_REG * NMS_REG);  // max_mask
    uint8_t* P = M + NMS_REG * NMS_REG;                               // supp_mask
    const int b = blockIdx.z;
    const int y0 = blockIdx.y * NMS_TILE - NMS_HALO, x0 = blockIdx.x * NMS_TILE - NMS_HALO;
    const float* sc = score + (size_t)b * h * w;
    const float NEG = -INFINITY;
    for (int i = threadIdx.x; i < NMS_REG * NMS_REG; i += blockDim.x) {
        const int y = y0 + i / NMS_REG, x = x0 + i % NMS_REG;
        S[i] = (y >= 0 && y < h && x >= 0 && x < w) ? sc[(size_t)y * w + x] : NEG;
    }
    __syncthreads();
    pool_rows(S, T, radius);
    __syncthreads();
    pool_cols(T, A, radius);
    __syncthreads();
    for (int i = threadIdx.x; i < NMS_REG * NMS_REG; i += blockDim.x) {
        const bool inside = S[i] != NEG;
        M[i] = (inside && S[i] == A[i]) ? 1 : 0;
    }
    __syncthreads();
    for (int it = 0; it < 2; ++it) {
        for (int i = threadIdx.x; i < NMS_REG * NMS_REG; i += blockDim.x) F[i] = M[i] ? 1.f : 0.f;
        __syncthreads();
        pool_rows(F, T, radius);
        __syncthreads();
        pool_cols(T, A, radius);
        __syncthreads();
        for (int i = threadIdx.x; i < NMS_REG * NMS_REG; i += blockDim.x) {
            const bool inside = S[i] != NEG;
            const bool supp = A[i] > 0.f;
            P[i] = supp ? 1 : 0;
            F[i] = inside ? (supp ? 0.f : S[i]) : NEG;  // supp_scores
        }
        __syncthreads();
        pool_rows(F, T, radius);
        __syncthreads();
        pool_cols(T, A, radius);
        __syncthreads();
        for (int i = threadIdx.x; i < NMS_REG * NMS_REG; i += blockDim.x) {
            const bool inside = S[i] != NEG;
            const bool newmax = inside && (F[i] == A[i]);
            if (newmax && !P[i]) M[i] = 1;
        }
        __syncthreads();
    }
    // emit survivors of the central tile
    for (int i = threadIdx.x; i < NMS_TILE * NMS_TILE; i += blockDim.x) {
        const int ly = i / NMS_TILE + NMS_HALO, lx = i % NMS_TILE + NMS_HALO;
        const int y = y0 + ly, x = x0 + lx;
        const int j = ly * NMS_REG + lx;
        bool keep = false;
        float s = 0.f;
        if (y < h && x < w) {
            s = S[j];
            keep = M[j] && s > threshold && y >= border && y < h - border && x >= border && x < w - border;
        }
        const unsigned ballot = __ballot_sync(__activemask(), keep);
        if (keep) {
            const int lane = threadIdx.x & 31;
            const int leader = __ffs(ballot) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&cand_count[slot0 + b], __popc(ballot));
            base = __shfl_sync(ballot, base, leader);
            const int pos = base + __popc(ballot & ((1u << lane) - 1));
            if (pos < GNB_CAND_CAP) {
                const unsigned long long key =
                    ((unsigned long long)(0xFFFFFFFFu - __float_as_uint(s)) << 32) | (unsigned)(y * w + x);
                cand_keys[(size_t)(slot0 + b) * GNB_CAND_CAP + pos] = key;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// NMS specialised for radius 4 (the SuperPoint default): 64x64 output tile + 20 px halo = 104x104
// region in shared memory.
//   * the three score max-pools are separable 9-wide filters; a work item owns a run of R consecutive
//     outputs of one row/column, loads its R+8 inputs once and forms the maxima by doubling
//     (m2 -> m4 -> m8 -> m9): 5 max ops per output.  Each pool only covers the region the later
//     stages still need (96^2, 80^2, 64^2 outputs), and "suppressed -> 0" is applied on the fly from
//     the suppression bit mask instead of materialising supp_scores;
//   * the two mask max-pools are dilations: done on 128-bit row bitmasks with shifts and ORs;
//   * adjacent lanes own adjacent lines => row pass strides by the (odd) pitch, column pass is
//     contiguous: no bank conflicts.
// Compare-only arithmetic, same decisions as simple_nms => bit-identical survivors.
#define N4_TILE 64
#define N4_REG 104
#define N4_PITCH 105
#define N4_THREADS 416
#define N4_WORDS 4   // 104 columns -> 4 x 32-bit mask words per row

// one separable pass; outputs [out_lo, out_lo + 8*R) along the run direction for lines [line_lo, line_lo + n_lines)
template <bool ROWS, int R, typename Src, typename Sink>
__device__ __forceinline__ void n4_pass(int line_lo, int n_lines, int out_lo, Src src, Sink sink) {
    const float NEG = -INFINITY;
    for (int item = threadIdx.x; item < n_lines * 8; item += N4_THREADS) {
        const int line = line_lo + item % n_lines, run = item / n_lines;
        const int o0 = out_lo + run * R, p0 = o0 - 4;
        float v[R + 8];
        src(line, p0, v);   // v[k] = input at position p0 + k along the run (NEG outside the region)
        float m2[R + 7], m4[R + 5], m8[R + 1];
#pragma unroll
        for (int k = 0; k < R + 7; ++k) m2[k] = fmaxf(v[k], v[k + 1]);
#pragma unroll
        for (int k = 0; k < R + 5; ++k) m4[k] = fmaxf(m2[k], m2[k + 2]);
#pragma unroll
        for (int k = 0; k < R + 1; ++k) m8[k] = fmaxf(m4[k], m4[k + 4]);
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const float m9 = fmaxf(m8[k], v[k + 8]);
            if (ROWS) sink(line, o0 + k, m9); else sink(o0 + k, line, m9);
        }
    }
    (void)NEG;
}

// 9x9 dilation of the bit mask `in` into `out` (rows of 4 words); `tmp` holds the horizontal pass
__device__ __forceinline__ void n4_dilate(const unsigned* __restrict__ in, unsigned* __restrict__ tmp, unsigned* __restrict__ out) {
    for (int y = threadIdx.x; y < N4_REG; y += N4_THREADS) {
        unsigned r[4], a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = in[y * N4_WORDS + i];
        auto shl = [](const unsigned x[4], int s, unsigned o[4]) {   // towards higher column index
            o[0] = x[0] << s;
#pragma unroll
            for (int i = 1; i < 4; ++i) o[i] = (x[i] << s) | (x[i - 1] >> (32 - s));
        };
        shl(r, 1, a);
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] |= r[i];            // offsets 0..1
        shl(a, 2, b);
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] |= b[i];            // 0..3
        shl(a, 4, b);
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] |= b[i];            // 0..7
        shl(r, 8, b);
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] |= b[i];            // 0..8
        // centre: shift right by 4 (towards lower column index)
#pragma unroll
        for (int i = 0; i < 4; ++i) tmp[y * N4_WORDS + i] = (a[i] >> 4) | (i < 3 ? (a[i + 1] << 28) : 0u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N4_REG * N4_WORDS; i += N4_THREADS) {
        const int y = i / N4_WORDS, wdx = i % N4_WORDS;
        const int lo = max(y - 4, 0), hi = min(y + 4, N4_REG - 1);
        unsigned acc = 0;
        for (int yy = lo; yy <= hi; ++yy) acc |= tmp[yy * N4_WORDS + wdx];
        out[i] = acc;
    }
}

__global__ void __launch_bounds__(N4_THREADS, 2) nms_r4_kernel(const float* __restrict__ score, int h, int w, float threshold,
                                                               int border, int slot0, unsigned long long* __restrict__ cand_keys,
                                                               int* __restrict__ cand_count) {
    extern __shared__ float sm[];
    float* S = sm;                                   // scores, -inf outside the image
    float* T = S + N4_REG * N4_PITCH;                // row-pass result
    unsigned* M = reinterpret_cast<unsigned*>(T + N4_REG * N4_PITCH);   // max_mask bits
    unsigned* P = M + N4_REG * N4_WORDS;             // supp_mask bits
    unsigned* H = P + N4_REG * N4_WORDS;             // dilation temp
    const int b = blockIdx.z;
    const int y0 = blockIdx.y * N4_TILE - NMS_HALO, x0 = blockIdx.x * N4_TILE - NMS_HALO;
    const float* sc = score + (size_t)b * h * w;
    const float NEG = -INFINITY;
    for (int i = threadIdx.x; i < N4_REG * N4_REG; i += N4_THREADS) {
        const int ly = i / N4_REG, lx = i - ly * N4_REG;
        const int y = y0 + ly, x = x0 + lx;
        S[ly * N4_PITCH + lx] = (y >= 0 && y < h && x >= 0 && x < w) ? __ldg(&sc[(size_t)y * w + x]) : NEG;
    }
    for (int i = threadIdx.x; i < N4_REG * N4_WORDS; i += N4_THREADS) M[i] = 0u;
    __syncthreads();

    auto in_range = [](int p) { return p >= 0 && p < N4_REG; };
    // sources: plain scores / scores with suppressed pixels set to 0 / row-pass result along columns
    auto srcS_row = [&](int y, int p0, float* v) {
#pragma unroll
        for (int k = 0; k < 20; ++k) v[k] = in_range(p0 + k) ? S[y * N4_PITCH + p0 + k] : NEG;
    };
    auto setM = [&](int y, int x) { atomicOr(&M[y * N4_WORDS + (x >> 5)], 1u << (x & 31)); };
    auto getP = [&](int y, int x) { return (P[y * N4_WORDS + (x >> 5)] >> (x & 31)) & 1u; };

    // pool 1 on raw scores: outputs [4,100)^2, runs of 12
    n4_pass<true, 12>(0, N4_REG, 4, srcS_row, [&](int y, int x, float m) { T[y * N4_PITCH + x] = m; });
    __syncthreads();
    n4_pass<false, 12>(4, 96, 4,
        [&](int x, int p0, float* v) {
#pragma unroll
            for (int k = 0; k < 20; ++k) v[k] = in_range(p0 + k) ? T[(p0 + k) * N4_PITCH + x] : NEG;
        },
        [&](int y, int x, float m) {
            const float s = S[y * N4_PITCH + x];
            if (s != NEG && s == m) setM(y, x);
        });
    __syncthreads();

    // iteration 1: supp valid on [8,96)^2, pooled suppressed scores on [12,92)^2 (runs of 10)
    n4_dilate(M, H, P);
    __syncthreads();
    n4_pass<true, 10>(8, 88, 12,
        [&](int y, int p0, float* v) {
#pragma unroll
            for (int k = 0; k < 18; ++k) {
                const int x = p0 + k;
                const float s = S[y * N4_PITCH + x];
                v[k] = (s != NEG && getP(y, x)) ? 0.f : s;
            }
        },
        [&](int y, int x, float m) { T[y * N4_PITCH + x] = m; });
    __syncthreads();
    n4_pass<false, 10>(12, 80, 12,
        [&](int x, int p0, float* v) {
#pragma unroll
            for (int k = 0; k < 18; ++k) v[k] = T[(p0 + k) * N4_PITCH + x];
        },
        [&](int y, int x, float m) {
            const float s = S[y * N4_PITCH + x];
            if (s != NEG && s == m && !getP(y, x)) setM(y, x);   // new_max & ~supp  (supp_score == score when !supp)
        });
    __syncthreads();

    // iteration 2: supp valid on [16,88)^2, pooled on [20,84)^2 = the output tile (runs of 8)
    n4_dilate(M, H, P);
    __syncthreads();
    n4_pass<true, 8>(16, 72, 20,
        [&](int y, int p0, float* v) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int x = p0 + k;
                const float s = S[y * N4_PITCH + x];
                v[k] = (s != NEG && getP(y, x)) ? 0.f : s;
            }
        },
        [&](int y, int x, float m) { T[y * N4_PITCH + x] = m; });
    __syncthreads();
    n4_pass<false, 8>(20, 64, 20,
        [&](int x, int p0, float* v) {
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = T[(p0 + k) * N4_PITCH + x];
        },
        [&](int y, int x, float m) {
            const float s = S[y * N4_PITCH + x];
            if (s != NEG && s == m && !getP(y, x)) setM(y, x);
        });
    __syncthreads();

    for (int i = threadIdx.x; i < N4_TILE * N4_TILE; i += N4_THREADS) {
        const int ly = i / N4_TILE + NMS_HALO, lx = i % N4_TILE + NMS_HALO;
        const int y = y0 + ly, x = x0 + lx;
        bool keep = false;
        float s = 0.f;
        if (y < h && x < w) {
            s = S[ly * N4_PITCH + lx];
            keep = ((M[ly * N4_WORDS + (lx >> 5)] >> (lx & 31)) & 1u) && s > threshold && y >= border && y < h - border &&
                   x >= border && x < w - border;
        }
        const unsigned ballot = __ballot_sync(__activemask(), keep);
        if (keep) {
            const int lane = threadIdx.x & 31;
            const int leader = __ffs(ballot) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&cand_count[slot0 + b], __popc(ballot));
            base = __shfl_sync(ballot, base, leader);
            const int pos = base + __popc(ballot & ((1u << lane) - 1));
            if (pos < GNB_CAND_CAP) {
                const unsigned long long key =
                    ((unsigned long long)(0xFFFFFFFFu - __float_as_uint(s)) << 32) | (unsigned)(y * w + x);
                cand_keys[(size_t)(slot0 + b) * GNB_CAND_CAP + pos] = key;
            }
        }
    }
}

// One CTA per image: exact top-K of the candidate keys (radix select on the 64-bit key), then a
// bitonic sort of the K selected keys in shared memory.
__global__ void __launch_bounds__(1024) topk_kernel(const unsigned long long* __restrict__ cand_keys,
                                                    const int* __restrict__ cand_count, int slot0, int w, int k_cap,
                                                    float* __restrict__ kp_xy, float* __restrict__ kp_score,
                                                    int* __restrict__ kp_count) {
    __shared__ unsigned long long sel[GNB_MAX_KP];
    __shared__ int hist[256];
    __shared__ unsigned long long s_prefix, s_mask;
    __shared__ int s_remaining, s_nsel;
    const int slot = slot0 + blockIdx.x;
    const int n_raw = cand_count[slot];
    if (n_raw > GNB_CAND_CAP) {  // overflow: candidates were dropped in atomic order -> refuse
        if (threadIdx.x == 0) kp_count[slot] = -1;
        return;
    }
    const int n = n_raw;
    const unsigned long long* keys = cand_keys + (size_t)slot * GNB_CAND_CAP;
    unsigned long long kth = ~0ull;
    if (n > k_cap) {
        if (threadIdx.x == 0) { s_prefix = 0; s_mask = 0; s_remaining = k_cap; }
        __syncthreads();
        for (int pass = 7; pass >= 0; --pass) {
            for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
            __syncthreads();
            const unsigned long long prefix = s_prefix, mask = s_mask;
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const unsigned long long key = keys[i];
                if ((key & mask) == prefix) atomicAdd(&hist[(int)((key >> (8 * pass)) & 255ull)], 1);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int cum = 0, d = 0;
                for (; d < 255; ++d) {
                    if (cum + hist[d] >= s_remaining) break;
                    cum += hist[d];
                }
                s_remaining -= cum;
                s_prefix |= (unsigned long long)d << (8 * pass);
                s_mask |= 255ull << (8 * pass);
            }
            __syncthreads();
        }
        kth = s_prefix;
    }
    if (threadIdx.x == 0) s_nsel = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long key = keys[i];
        if (key <= kth) {
            const int pos = atomicAdd(&s_nsel, 1);
            if (pos < GNB_MAX_KP) sel[pos] = key;
        }
    }
    __syncthreads();
    const int nsel = min(s_nsel, k_cap);
    int npow = 1;
    while (npow < nsel) npow <<= 1;
    for (int i = nsel + threadIdx.x; i < npow; i += blockDim.x) sel[i] = ~0ull;
    __syncthreads();
    for (int size = 2; size <= npow; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < npow; i += blockDim.x) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool up = (i & size) == 0;
                    const unsigned long long a = sel[i], c = sel[j];
                    if ((a > c) == up) { sel[i] = c; sel[j] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < nsel; i += blockDim.x) {
        const unsigned long long key = sel[i];
        const unsigned idx = (unsigned)(key & 0xFFFFFFFFull);
        const float s = __uint_as_float(0xFFFFFFFFu - (unsigned)(key >> 32));
        kp_xy[((size_t)slot * k_cap + i) * 2 + 0] = (float)(idx % (unsigned)w);
        kp_xy[((size_t)slot * k_cap + i) * 2 + 1] = (float)(idx / (unsigned)w);
        kp_score[(size_t)slot * k_cap + i] = s;
    }
    if (threadIdx.x == 0) kp_count[slot] = nsel;
}

int gnb_kp_select(gnb_ctx* ctx, const float* score, int n, int h, int w, int slot0) {
    if (ctx->cfg.nms_radius * 5 > NMS_HALO) {
        GNB_SET_ERR(ctx, "nms_radius %d too large for the NMS halo", ctx->cfg.nms_radius);
        return GNB_E_INVALID;
    }
    GNB_CUDA(ctx, cudaMemsetAsync(ctx->cand_count + slot0, 0, sizeof(int) * n, ctx->stream));
    if (ctx->cfg.nms_radius == 4) {
        const size_t smem = (size_t)N4_REG * N4_PITCH * 2 * sizeof(float) + 3 * N4_REG * N4_WORDS * sizeof(unsigned);
        GNB_CUDA(ctx, gnb_func_smem(ctx, nms_r4_kernel, (int)smem));
        dim3 grid(ceil_div(w, N4_TILE), ceil_div(h, N4_TILE), n);
        GNB_KERNEL(ctx, "nms_r4_kernel", nms_r4_kernel<<<grid, N4_THREADS, smem, ctx->stream>>>(
            score, h, w, ctx->cfg.keypoint_threshold, ctx->cfg.border, slot0, ctx->cand_keys, ctx->cand_count));
    } else {
        const size_t smem = (size_t)NMS_REG * NMS_REG * (4 * sizeof(float) + 2);
        GNB_CUDA(ctx, gnb_func_smem(ctx, nms_kernel, (int)smem));
        dim3 grid(ceil_div(w, NMS_TILE), ceil_div(h, NMS_TILE), n);
        GNB_KERNEL(ctx, "nms_kernel", nms_kernel<<<grid, 256, smem, ctx->stream>>>(score, h, w, ctx->cfg.nms_radius, ctx->cfg.keypoint_threshold,
                                                     ctx->cfg.border, slot0, ctx->cand_keys, ctx->cand_count));
    }
    GNB_KERNEL(ctx, "topk_kernel", topk_kernel<<<n, 1024, 0, ctx->stream>>>(ctx->cand_keys, ctx->cand_count, slot0, w, ctx->cfg.max_keypoints,
                                             ctx->kp_xy, ctx->kp_score, ctx->kp_count));
    return GNB_OK;
}

// ------------------------------------------------------------------------------------------------
// K3: one warp per keypoint; lane l owns channels [8l, 8l+8).  Reads 4 x 1 KB, writes 1 KB.
__global__ void __launch_bounds__(256) sample_kernel(const float* __restrict__ dense, int hc, int wc, int img_h,
                                                     int img_w, const float* __restrict__ kp_xy,
                                                     const int* __restrict__ kp_count, int slot0, int k_cap,
                                                     float* __restrict__ desc) {
    const int b = blockIdx.y;
    const int slot = slot0 + b;
    const int kp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (kp >= kp_count[slot]) return;
    const float x = kp_xy[((size_t)slot * k_cap + kp) * 2 + 0], y = kp_xy[((size_t)slot * k_cap + kp) * 2 + 1];
    // grid_sample(align_corners=True) on g = (kp - 3.5) / (dim - 4.5) * 2 - 1
    const float gx = __fdiv_rn(__fsub_rn(x, 3.5f), (float)img_w - 4.5f);
    const float gy = __fdiv_rn(__fsub_rn(y, 3.5f), (float)img_h - 4.5f);
    const float fx = __fmul_rn(gx, (float)(wc - 1)), fy = __fmul_rn(gy, (float)(hc - 1));
    const float x0f = floorf(fx), y0f = floorf(fy);
    const int x0 = (int)x0f, y0 = (int)y0f;
    const float ax = __fsub_rn(fx, x0f), ay = __fsub_rn(fy, y0f);
    const float w00 = __fmul_rn(1.f - ax, 1.f - ay), w01 = __fmul_rn(ax, 1.f - ay);
    const float w10 = __fmul_rn(1.f - ax, ay), w11 = __fmul_rn(ax, ay);
    const float* base = dense + (size_t)b * hc * wc * 256;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    auto tap = [&](int yy, int xx, float wgt) {
        if (yy < 0 || yy >= hc || xx < 0 || xx >= wc) return;
        const float4* p = reinterpret_cast<const float4*>(base + ((size_t)yy * wc + xx) * 256 + lane * 8);
        const float4 a = __ldg(p), c = __ldg(p + 1);
        v[0] = __fadd_rn(v[0], __fmul_rn(a.x, wgt)); v[1] = __fadd_rn(v[1], __fmul_rn(a.y, wgt));
        v[2] = __fadd_rn(v[2], __fmul_rn(a.z, wgt)); v[3] = __fadd_rn(v[3], __fmul_rn(a.w, wgt));
        v[4] = __fadd_rn(v[4], __fmul_rn(c.x, wgt)); v[5] = __fadd_rn(v[5], __fmul_rn(c.y, wgt));
        v[6] = __fadd_rn(v[6], __fmul_rn(c.z, wgt)); v[7] = __fadd_rn(v[7], __fmul_rn(c.w, wgt));
    };
    tap(y0, x0, w00); tap(y0, x0 + 1, w01); tap(y0 + 1, x0, w10); tap(y0 + 1, x0 + 1, w11);
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += v[j] * v[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float nrm = fmaxf(sqrtf(ss), 1e-12f);
    float4* o = reinterpret_cast<float4*>(desc + ((size_t)slot * k_cap + kp) * 256 + lane * 8);
    o[0] = make_float4(v[0] / nrm, v[1] / nrm, v[2] / nrm, v[3] / nrm);
    o[1] = make_float4(v[4] / nrm, v[5] / nrm, v[6] / nrm, v[7] / nrm);
}

int gnb_kp_sample(gnb_ctx* ctx, const float* dense, int n, int h, int w, int slot0) {
    const int k = ctx->cfg.max_keypoints;
    dim3 grid(ceil_div(k * 32, 256), n);
    GNB_KERNEL(ctx, "sample_kernel", sample_kernel<<<grid, 256, 0, ctx->stream>>>(dense, h / 8, w / 8, h, w, ctx->kp_xy, ctx->kp_count, slot0, k,
                                                 ctx->desc_f32));
    return GNB_OK;
}
