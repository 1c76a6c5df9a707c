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
#include <stdlib.h>

static inline unsigned __float_as_uint_host(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

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
// Sparse, top-K-aware NMS (radius 4) — the product path.
//
// Lemma (tests/test_oracle_cpu.py::test_nms_survivors_above_a_level_depend_only_on_pixels_above_it): for any level T
// the survivors of simple_nms with score > T are unchanged when every pixel <= T is replaced by 0 — a pixel can only
// suppress pixels that are not larger than itself.  Only the K best survivors are kept, and with the shipped weights
// < 1 % (K = 1024) to 3 % (K = 2048) of the pixels score above the K-th survivor.  So:
//   1. nms_hist_kernel   histogram of the score bit patterns (every 8th row: the level only steers the work, never the
//                        result); its last CTA per image picks a level L with ~16 K pixels at or above it
//                        (never below the keypoint threshold);
//   2. nms_sparse_kernel per 64x64 tile + 20 px halo: scores below L enter shared memory as 0, the pixels at or above
//                        L form a short list, and every stage of simple_nms (window max, suppression dilation, two
//                        refinement rounds) touches only listed pixels: a 3x3 pre-check, then the full 9x9 window for
//                        the few that pass.  Compare-only arithmetic on the same values => the same decisions;
//   3. nms_check_kernel  an image with fewer than K survivors above its level (rare: flat score maps) is flagged, its
//                        level drops to the plain threshold and a second launch of nms_sparse_kernel (whose CTAs exit
//                        at once when nothing is flagged) redoes that image exactly.
// Bit-identical to oracle/nms_ref.py by construction; tests/test_gpu_parity.py::test_k2_* and test_gpu_fullsize.py.
#define NS_TILE 64
#define NS_REG 104
#define NS_PITCH 105   // odd pitch: the window scans walk columns
#define NS_WORDS 4
#define NS_THREADS 256
#define NS_BUCKETS 2048          // score bits >> 19: sign + exponent + 4 mantissa bits of a float in [0, 2)
#define NS_HIST_ROW_STEP 8

// The last CTA of an image to finish (ticket) turns the merged histogram into that image's level: the highest bucket
// edge with >= target sampled pixels at or above it, never below the keypoint threshold — and clears histogram and ticket
// for the next call.  (Plain shared-memory atomics: `__match_any_sync` aggregation was measured slower, here and in
// `topk_kernel`.)
__global__ void __launch_bounds__(256) nms_hist_kernel(const float* __restrict__ score, int h, int w, unsigned thr_bits, int slot0,
                                                       unsigned* __restrict__ hist, unsigned* __restrict__ ticket, unsigned target_sampled,
                                                       unsigned* __restrict__ level, int* __restrict__ flag, int* __restrict__ cand_count,
                                                       int* __restrict__ list_count) {
    __shared__ unsigned sh[NS_BUCKETS];
    __shared__ unsigned s_warp[8];
    __shared__ unsigned s_edge;
    __shared__ int s_last;
    for (int i = threadIdx.x; i < NS_BUCKETS; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    const int b = blockIdx.z, slot = slot0 + b;
    const float* sc = score + (size_t)b * h * w;
    const int rows = (h + NS_HIST_ROW_STEP - 1) / NS_HIST_ROW_STEP;
    const int total = rows * w;
    const int stride = gridDim.x * blockDim.x;
    for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
        unsigned bits[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {     // four independent loads in flight
            const int i = i0 + u * stride;
            bits[u] = 0u;
            if (i < total) {
                const int r = i / w, x = i - r * w;
                bits[u] = __float_as_uint(__ldg(&sc[(size_t)(r * NS_HIST_ROW_STEP) * w + x]));
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if ((int)bits[u] > (int)thr_bits && bits[u] < 0x40000000u) atomicAdd(&sh[bits[u] >> 19], 1u);   // thr < score < 2
    }
    __syncthreads();
    unsigned* hs = hist + (size_t)slot * NS_BUCKETS;
    for (int i = threadIdx.x; i < NS_BUCKETS; i += blockDim.x)
        if (sh[i]) atomicAdd(&hs[i], sh[i]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&ticket[slot], 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- level: thread t owns buckets [8 t, 8 t + 8); suffix sums over threads (thread 255 holds the highest buckets)
    constexpr int PER = NS_BUCKETS / 256;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned cnt[PER], mine = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) { cnt[i] = __ldcg(&hs[threadIdx.x * PER + i]); mine += cnt[i]; }
    unsigned incl = mine;   // inclusive suffix sum inside the warp (lanes above this one)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += t;
    }
    if (lane == 0) s_warp[warp] = incl;
    if (threadIdx.x == 0) s_edge = 0u;
    __syncthreads();
    unsigned above = incl - mine;   // pixels in buckets strictly above this thread's
    for (int wi = warp + 1; wi < 8; ++wi) above += s_warp[wi];
    if (above < target_sampled && above + mine >= target_sampled) {   // the crossing thread (at most one)
        unsigned acc = above;
        for (int i = PER - 1; i >= 0; --i) {
            acc += cnt[i];
            if (acc >= target_sampled) { s_edge = (unsigned)(threadIdx.x * PER + i); break; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned lo = thr_bits + 1u;                    // score > threshold  <=>  bits >= thr_bits + 1 (non-negative floats)
        level[slot] = max(s_edge << 19, lo);                  // no crossing (fewer pixels than the target): the threshold itself
        flag[slot] = 0;
        ticket[slot] = 0u;
        cand_count[slot] = 0;      // the per-call counters restart here instead of in two memset launches
        list_count[b] = 0;
    }
#pragma unroll
    for (int i = 0; i < PER; ++i) hs[threadIdx.x * PER + i] = 0u;   // ready for the next call
}

// first_count[slot] = survivors of the first pass; the counters of the images about to be redone restart at 0
__global__ void nms_redo_prepare(int n, int slot0, int k_cap, unsigned thr_bits, const unsigned* __restrict__ level, int* __restrict__ cand_count,
                                 int* __restrict__ first_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int slot = slot0 + i;
    const int c = cand_count[slot];
    first_count[slot] = c;
    if (c < k_cap && level[slot] > thr_bits + 1u) cand_count[slot] = 0;
}

#define NS_LIST_CAP 6144   // listed pixels per region kept as a list; a denser region is walked pixel by pixel instead
#define NS_SURV_CAP 4096   // survivors of one 64x64 tile (every pixel of a tied plateau survives)

__device__ __forceinline__ unsigned ns_getbit(const unsigned* m, int y, int x) { return (m[y * NS_WORDS + (x >> 5)] >> (x & 31)) & 1u; }

struct NsRegion {
    const float* S;            // [NS_REG][NS_PITCH] scores at or above the level, 0 elsewhere
    unsigned* MAXM;            // max_mask
    unsigned* SUPA;            // supp_mask after the initial maxima       = dilate(max_mask_0)
    unsigned* SUPB;            // added by the maxima of refinement round 1 = dilate(new maxima of round 1)
    const unsigned short* list;
    unsigned short* surv;
    int* n_surv;
    int n;                     // listed pixels (> NS_LIST_CAP: walk the region densely)
};

// One pass over the listed pixels of a region.  Each lane ring-checks its own pixel (3x3: no larger unsuppressed direct
// neighbour), then the WARP scans the 9x9 window of every pixel that passed (81 positions over 32 lanes, one vote).
// A new maximum sets its max_mask bit, ORs its 9x9 block into `sup_out` (the dilation simple_nms applies to the mask,
// built incrementally) and, inside the output tile, joins the survivor list.
// ROUND 0: plain window max.  ROUND 1: suppressed = SUPA.  ROUND 2: suppressed = SUPA | SUPB.
template <int ROUND>
__device__ __forceinline__ void ns_mark_maxima(const NsRegion& R, unsigned* __restrict__ sup_out, int lo, int hi) {
    const int lane = threadIdx.x & 31;
    const bool dense = R.n > NS_LIST_CAP;
    const int count = dense ? NS_REG * NS_REG : R.n;
    auto suppressed = [&](int y, int x) -> bool {
        if (ROUND == 0) return false;
        unsigned wv = R.SUPA[y * NS_WORDS + (x >> 5)];
        if (ROUND == 2) wv |= R.SUPB[y * NS_WORDS + (x >> 5)];
        return (wv >> (x & 31)) & 1u;
    };
    for (int e0 = (threadIdx.x >> 5) * 32; e0 < count; e0 += NS_THREADS) {
        const int e = e0 + lane;
        int ly = 0, lx = 0;
        float s = 0.f;
        bool cand = false;
        if (e < count) {
            if (dense) { ly = e / NS_REG; lx = e - ly * NS_REG; } else { ly = R.list[e] >> 8; lx = R.list[e] & 255; }
            if (ly >= lo && ly < hi && lx >= lo && lx < hi) {
                s = R.S[ly * NS_PITCH + lx];
                cand = s > 0.f && !suppressed(ly, lx);
                if (cand) {
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                        for (int dx = -1; dx <= 1; ++dx)
                            if ((dy | dx) != 0 && R.S[(ly + dy) * NS_PITCH + lx + dx] > s && !suppressed(ly + dy, lx + dx)) cand = false;
                }
            }
        }
        unsigned todo = __ballot_sync(0xffffffffu, cand);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1u;
            const int cy = __shfl_sync(0xffffffffu, ly, src), cx = __shfl_sync(0xffffffffu, lx, src);
            const float cs = __shfl_sync(0xffffffffu, s, src);
            bool bad = false;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int p = lane + 32 * r;   // window position 0..80
                if (p < 81) {
                    const int yy = cy + p / 9 - 4, xx = cx + p % 9 - 4;
                    if (yy >= 0 && yy < NS_REG && xx >= 0 && xx < NS_REG && R.S[yy * NS_PITCH + xx] > cs && !suppressed(yy, xx)) bad = true;
                }
            }
            if (__any_sync(0xffffffffu, bad)) continue;
            if (lane == 0) {
                atomicOr(&R.MAXM[cy * NS_WORDS + (cx >> 5)], 1u << (cx & 31));
                if (cy >= 20 && cy < 20 + NS_TILE && cx >= 20 && cx < 20 + NS_TILE) {
                    const int pos = atomicAdd(R.n_surv, 1);
                    if (pos < NS_SURV_CAP) R.surv[pos] = (unsigned short)((cy << 8) | cx);
                }
            }
            if (sup_out && lane < 9) {   // rows cy - 4 .. cy + 4, columns cx - 4 .. cx + 4 (clipped to the region)
                const int yy = cy - 4 + lane;
                if (yy >= 0 && yy < NS_REG) {
                    const int x_lo = max(cx - 4, 0), x_hi = min(cx + 4, NS_REG - 1);
                    const int w0 = x_lo >> 5, w1 = x_hi >> 5;
                    const unsigned m_lo = 0xffffffffu << (x_lo & 31), m_hi = 0xffffffffu >> (31 - (x_hi & 31));
                    if (w0 == w1) atomicOr(&sup_out[yy * NS_WORDS + w0], m_lo & m_hi);
                    else { atomicOr(&sup_out[yy * NS_WORDS + w0], m_lo); atomicOr(&sup_out[yy * NS_WORDS + w1], m_hi); }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(NS_THREADS, 3) nms_sparse_kernel(const float* __restrict__ score, int h, int w, int n_img, float threshold, int border,
                                                                int slot0, const unsigned* __restrict__ level, const int* __restrict__ first_count,
                                                                int redo, int k_cap, unsigned thr_bits, unsigned long long* __restrict__ cand_keys,
                                                                int* __restrict__ cand_count, const int* __restrict__ any_redo = nullptr) {
    if (any_redo && *any_redo == 0) return;
    extern __shared__ float sm[];
    float* S = sm;
    unsigned* MAXM = reinterpret_cast<unsigned*>(S + NS_REG * NS_PITCH);
    unsigned* SUPA = MAXM + NS_REG * NS_WORDS;
    unsigned* SUPB = SUPA + NS_REG * NS_WORDS;
    unsigned short* list = reinterpret_cast<unsigned short*>(SUPB + NS_REG * NS_WORDS);   // [NS_LIST_CAP] (ly << 8 | lx)
    unsigned short* surv = list + NS_LIST_CAP;                                            // [NS_SURV_CAP]
    __shared__ int n_act, n_surv;
    const int tiles_x = (w + NS_TILE - 1) / NS_TILE, tiles_y = (h + NS_TILE - 1) / NS_TILE;
    const int tiles = tiles_x * tiles_y, items = tiles * n_img;
    const int lane = threadIdx.x & 31;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int b = item / tiles, t = item - b * tiles;
        const int slot = slot0 + b;
        unsigned lv = level[slot];
        if (redo) {
            // second launch: only images whose first pass found fewer than K survivors above a level that was higher than
            // the plain threshold are redone, from the threshold (first_count / level are not written by this launch)
            // (redo == 2, list path: first_count holds the decision itself, see nms_redo_prepare_lists)
            if (redo == 2 ? first_count[slot] == 0 : !(first_count[slot] < k_cap && lv > thr_bits + 1u)) continue;
            lv = thr_bits + 1u;
        }
        const int y0 = (t / tiles_x) * NS_TILE - 20, x0 = (t % tiles_x) * NS_TILE - 20;
        const float* sc = score + (size_t)b * h * w;
        __syncthreads();   // previous item's shared memory is no longer read
        if (threadIdx.x == 0) { n_act = 0; n_surv = 0; }
        for (int i = threadIdx.x; i < 3 * NS_REG * NS_WORDS; i += NS_THREADS) MAXM[i] = 0u;   // MAXM, SUPA, SUPB are contiguous
        __syncthreads();
        // ---- load.  Fast path (row pitch a multiple of 4 pixels): the region row is 26 aligned float4 (x0 = 64 t - 20 is a
        // multiple of 4, so a vector is entirely inside or outside the image); every thread issues all of its loads
        // before the first use, then compacts its listed pixels with one warp scan + one atomicAdd per warp.
        if ((w & 3) == 0) {
            constexpr int V = NS_REG / 4;                                   // 26 vectors per row
            constexpr int PER = (NS_REG * V + NS_THREADS - 1) / NS_THREADS;  // 11
            float4 v[PER];
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const int idx = u * NS_THREADS + threadIdx.x;
                const int ly = idx / V, c4 = idx - ly * V;
                const int y = y0 + ly, x = x0 + 4 * c4;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < NS_REG * V && y >= 0 && y < h && x >= 0 && x < w) v[u] = __ldg(reinterpret_cast<const float4*>(&sc[(size_t)y * w + x]));
            }
            unsigned long long bits = 0ull;
            int cnt = 0;
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const int idx = u * NS_THREADS + threadIdx.x;
                if (idx < NS_REG * V) {
                    const int ly = idx / V, c4 = idx - ly * V;
                    float f[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool act = (int)__float_as_uint(f[j]) >= (int)lv;
                        if (act) { bits |= 1ull << (4 * u + j); ++cnt; } else f[j] = 0.f;
                    }
                    float* dst = S + ly * NS_PITCH + 4 * c4;
                    dst[0] = f[0]; dst[1] = f[1]; dst[2] = f[2]; dst[3] = f[3];
                }
            }
            int incl = cnt;   // inclusive warp scan of the per-thread counts
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t2 = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t2;
            }
            const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
            int base = 0;
            if (lane == 31 && warp_total) base = atomicAdd(&n_act, warp_total);
            base = __shfl_sync(0xffffffffu, base, 31) + incl - cnt;
            while (bits) {
                const int bpos = __ffsll((long long)bits) - 1;
                bits &= bits - 1ull;
                const int idx = (bpos >> 2) * NS_THREADS + threadIdx.x;
                const int ly = idx / V, c4 = idx - ly * V;
                if (base < NS_LIST_CAP) list[base] = (unsigned short)((ly << 8) | (4 * c4 + (bpos & 3)));
                ++base;
            }
        } else {
            for (int i = threadIdx.x; i < NS_REG * 128; i += NS_THREADS) {
                const int ly = i >> 7, lx = i & 127;
                const int y = y0 + ly, x = x0 + lx;
                float v = 0.f;
                bool act = false;
                if (lx < NS_REG && y >= 0 && y < h && x >= 0 && x < w) {
                    v = __ldg(&sc[(size_t)y * w + x]);
                    act = (int)__float_as_uint(v) >= (int)lv;
                }
                if (lx < NS_REG) S[ly * NS_PITCH + lx] = act ? v : 0.f;
                const unsigned ballot = __ballot_sync(0xffffffffu, act);
                if (ballot) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&n_act, __popc(ballot));
                    base = __shfl_sync(0xffffffffu, base, 0) + __popc(ballot & ((1u << lane) - 1));
                    if (act && base < NS_LIST_CAP) list[base] = (unsigned short)((ly << 8) | lx);
                }
            }
        }
        __syncthreads();
        NsRegion R{S, MAXM, SUPA, SUPB, list, surv, &n_surv, n_act};
        if (R.n == 0) continue;   // uniform: nothing at or above the level in this region
        // ---- simple_nms: max_mask = (score == 9x9 max) on [4, 100)^2, then two refinement rounds, each valid on a region
        // 8 px smaller; the supp masks are the dilations of the maxima found so far, built by the finder itself.  A pixel of
        // the (scores > 0) list can only lose to a LARGER unsuppressed pixel, so "ties survive" exactly as == does upstream.
        ns_mark_maxima<0>(R, SUPA, 4, NS_REG - 4);
        __syncthreads();
        ns_mark_maxima<1>(R, SUPB, 12, NS_REG - 12);
        __syncthreads();
        ns_mark_maxima<2>(R, nullptr, 20, NS_REG - 20);
        __syncthreads();
        // ---- emit the survivors of the output tile (a maximum stays one: each was listed when it was found)
        const int ns = n_surv;
        if (ns <= NS_SURV_CAP) {
            if (threadIdx.x < 32 && ns > 0) {
                for (int e0 = 0; e0 < ns; e0 += 32) {
                    const int e = e0 + lane;
                    bool keep = false;
                    float sv = 0.f;
                    int y = 0, x = 0;
                    if (e < ns) {
                        const int ly = surv[e] >> 8, lx = surv[e] & 255;
                        y = y0 + ly; x = x0 + lx;
                        sv = S[ly * NS_PITCH + lx];
                        keep = sv > threshold && y >= border && y < h - border && x >= border && x < w - border;
                    }
                    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
                    if (ballot) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd(&cand_count[slot], __popc(ballot));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        const int pos = base + __popc(ballot & ((1u << lane) - 1));
                        if (keep && pos < GNB_CAND_CAP)
                            cand_keys[(size_t)slot * GNB_CAND_CAP + pos] = ((unsigned long long)(0xFFFFFFFFu - __float_as_uint(sv)) << 32) | (unsigned)(y * w + x);
                    }
                }
            }
        } else {
            // survivor list overflow (a large tied plateau): walk the tile's max_mask instead
            for (int i0 = 0; i0 < NS_TILE * NS_TILE; i0 += NS_THREADS) {
                const int i = i0 + threadIdx.x;
                const int ly = i / NS_TILE + 20, lx = i % NS_TILE + 20;
                const int y = y0 + ly, x = x0 + lx;
                const float sv = S[ly * NS_PITCH + lx];
                const bool keep = ns_getbit(MAXM, ly, lx) && sv > threshold && y >= border && y < h - border && x >= border && x < w - border;
                const unsigned ballot = __ballot_sync(0xffffffffu, keep);
                if (ballot) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&cand_count[slot], __popc(ballot));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    const int pos = base + __popc(ballot & ((1u << lane) - 1));
                    if (keep && pos < GNB_CAND_CAP)
                        cand_keys[(size_t)slot * GNB_CAND_CAP + pos] = ((unsigned long long)(0xFFFFFFFFu - __float_as_uint(sv)) << 32) | (unsigned)(y * w + x);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// List-based form of the same sparse NMS — the product path.  The tile kernel above spends its time in the three marking
// passes of 104x104 regions (62 % of its warp samples; 30 % of all samples wait at its barriers) and repeats the work of
// a halo pixel in up to four regions.  Here the listed pixels of the WHOLE image are compacted once and every stage of
// simple_nms is one launch over that list against the score map in L2 (59 MB for 16 frames: it stays resident):
//   nms_compact_kernel   streams the map once; (pixel index, score bits) of the pixels at or above the level -> nms_list
//   nms_round_kernel<R>  one lane per listed pixel: 3x3 pre-check, then the warp scans the 9x9 window of the pixels that
//                        passed.  A new maximum is final (it is emitted as a top-K candidate at once) and ORs its 9x9
//                        block into the suppression bitmap the NEXT round reads (R = 0 -> SUPA, 1 -> SUPB): no barrier,
//                        no halo, no shared memory.  Same comparisons on the same values as the tile kernel.
// An image whose list overflows (ties at the level, flat maps) or that ends with fewer than K survivors above its level
// is redone from the plain threshold by the tile kernel (nms_redo_prepare_lists -> nms_sparse_kernel(redo)).
#define NL_PER_THREAD 40   // pixels per thread of the compaction (ten float4 loads in flight)

// (A variant that kept the row-major pixel order inside the list — ballot ranks per step — made this kernel 2x slower,
// 0.21 vs 0.11 ms per 128 frames, and the round kernels no faster: they wait on load latency, not on L1 tag cycles.)
template <bool VEC>
__global__ void __launch_bounds__(256) nms_compact_kernel(const float* __restrict__ score, int hw, int slot0, const unsigned* __restrict__ level,
                                                          uint2* __restrict__ list, int* __restrict__ list_count) {
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lv = (int)level[slot0 + b];
    const float* sc = score + (size_t)b * hw;
    const int chunk0 = blockIdx.x * (256 * NL_PER_THREAD);
    unsigned vbits[NL_PER_THREAD];
    // element e of this thread is pixel pix(e): a float4 per thread and step (VEC) or one float (ragged sizes)
    auto pix = [&](int e) { return VEC ? chunk0 + ((e >> 2) * 256 + (int)threadIdx.x) * 4 + (e & 3) : chunk0 + e * 256 + (int)threadIdx.x; };
    if (VEC) {
#pragma unroll
        for (int u = 0; u < NL_PER_THREAD / 4; ++u) {
            const int i = pix(4 * u);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < hw) v = __ldg(reinterpret_cast<const float4*>(sc + i));   // hw % 4 == 0: a vector is inside or outside
            vbits[4 * u] = __float_as_uint(v.x); vbits[4 * u + 1] = __float_as_uint(v.y);
            vbits[4 * u + 2] = __float_as_uint(v.z); vbits[4 * u + 3] = __float_as_uint(v.w);
        }
    } else {
#pragma unroll
        for (int e = 0; e < NL_PER_THREAD; ++e) {
            const int i = pix(e);
            vbits[e] = i < hw ? __float_as_uint(__ldg(sc + i)) : 0u;
        }
    }
    // listed <=> bits >= level as signed ints (level >= 1: padding zeros and negative scores never count)
    int cnt = 0;
#pragma unroll
    for (int e = 0; e < NL_PER_THREAD; ++e) cnt += ((int)vbits[e] >= lv) ? 1 : 0;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int t = s_warp[i]; s_warp[i] = tot; tot += t; }
        s_base = tot ? atomicAdd(&list_count[b], tot) : 0;
    }
    __syncthreads();
    if (cnt == 0) return;
    int pos = s_base + s_warp[warp] + incl - cnt;
    uint2* out = list + (size_t)b * GNB_NMS_LIST_CAP;
#pragma unroll
    for (int e = 0; e < NL_PER_THREAD; ++e)
        if ((int)vbits[e] >= lv) {
            if (pos < GNB_NMS_LIST_CAP) out[pos] = make_uint2((unsigned)pix(e), vbits[e]);
            ++pos;
        }
}

// Latency is what this kernel fights (77 % of its warp samples waited on a load in the first version): every batch of
// loads is issued back to back from clamped addresses before the first compare, up to FOUR candidates of a warp-iteration
// are scanned at once (8 lanes each, 11 window positions per lane in flight), and the survivors of a warp are buffered in
// shared memory so that the candidate counter is bumped once per warp, not once per iteration.
// Bitmaps: round 0 ORs the 9x9 blocks of its maxima into SUPA and SUPB; round 1 reads SUPA and adds its own to SUPB;
// round 2 reads SUPB (= dilation of all maxima so far).  Each round reads one bitmap and never the one it writes.
#define NR_BUF 64   // buffered survivors per warp (flushed when more than half full)
template <int ROUND>
__global__ void __launch_bounds__(256) nms_round_kernel(const float* __restrict__ score, int h, int w, int slot0, const uint2* __restrict__ list,
                                                        const int* __restrict__ list_count, unsigned* __restrict__ sup_a,
                                                        unsigned* __restrict__ sup_b, int wpr, float threshold, int border,
                                                        unsigned long long* __restrict__ cand_keys, int* __restrict__ cand_count) {
    __shared__ unsigned long long s_keys[8][NR_BUF];
    const int b = blockIdx.y, slot = slot0 + b, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = list_count[b];
    if (n > GNB_NMS_LIST_CAP) return;   // overflow: this image is redone by the tile kernel
    const float* sc = score + (size_t)b * h * w;
    const uint2* li = list + (size_t)b * GNB_NMS_LIST_CAP;
    const size_t bm = (size_t)b * h * wpr;
    const unsigned* rd = (ROUND == 2 ? sup_b : sup_a) + bm;     // the bitmap this round reads (none in round 0)
    unsigned* wa = sup_a + bm;
    unsigned* wb = sup_b + bm;
    auto suppressed = [&](int y, int x) -> bool {
        if (ROUND == 0) return false;
        return (rd[y * wpr + (x >> 5)] >> (x & 31)) & 1u;
    };
    const int grp = lane >> 3, sub = lane & 7;
    int n_buf = 0;   // warp-uniform
    auto flush = [&]() {
        int base = 0;
        if (lane == 0) base = atomicAdd(&cand_count[slot], n_buf);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int i = lane; i < n_buf; i += 32)
            if (base + i < GNB_CAND_CAP) cand_keys[(size_t)slot * GNB_CAND_CAP + base + i] = s_keys[warp][i];
        __syncwarp();
        n_buf = 0;
    };
    for (int e0 = (blockIdx.x * 8 + warp) * 32; e0 < n; e0 += gridDim.x * 256) {
        const int e = e0 + lane;
        int py = 0, px = 0;
        float s = 0.f;
        bool cand = false;
        if (e < n) {
            const uint2 en = li[e];
            py = (int)en.x / w; px = (int)en.x - py * w;
            s = __uint_as_float(en.y);
            cand = !suppressed(py, px);    // rounds 1 and 2: most listed pixels sit inside the 9x9 block of an earlier maximum
        }
        if (cand) {
            // 3x3 pre-check: the eight loads leave together (clamped addresses: a clamped position is the pixel itself or
            // another member of its 3x3 block, so it changes nothing)
            const int ym = max(py - 1, 0), yp = min(py + 1, h - 1), xm = max(px - 1, 0), xp = min(px + 1, w - 1);
            const float* r0 = sc + (size_t)ym * w;
            const float* r1 = sc + (size_t)py * w;
            const float* r2 = sc + (size_t)yp * w;
            const float v0 = __ldg(r0 + xm), v1 = __ldg(r0 + px), v2 = __ldg(r0 + xp), v3 = __ldg(r1 + xm), v4 = __ldg(r1 + xp),
                        v5 = __ldg(r2 + xm), v6 = __ldg(r2 + px), v7 = __ldg(r2 + xp);
            if (v0 > s && !suppressed(ym, xm)) cand = false;
            if (v1 > s && !suppressed(ym, px)) cand = false;
            if (v2 > s && !suppressed(ym, xp)) cand = false;
            if (v3 > s && !suppressed(py, xm)) cand = false;
            if (v4 > s && !suppressed(py, xp)) cand = false;
            if (v5 > s && !suppressed(yp, xm)) cand = false;
            if (v6 > s && !suppressed(yp, px)) cand = false;
            if (v7 > s && !suppressed(yp, xp)) cand = false;
        }
        unsigned todo = __ballot_sync(0xffffffffu, cand);
        bool is_max = false;
        while (todo) {
            // up to four candidates at once: group g (8 lanes) takes the g-th set bit of `todo`
            int src = -1;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int f = todo ? __ffs(todo) - 1 : -1;
                if (todo) todo &= todo - 1u;
                if (g == grp) src = f;
            }
            const int sl = src < 0 ? 0 : src;
            const int cy = __shfl_sync(0xffffffffu, py, sl), cx = __shfl_sync(0xffffffffu, px, sl);
            const float cs = __shfl_sync(0xffffffffu, s, sl);
            float wv[11];
#pragma unroll
            for (int r = 0; r < 11; ++r) {
                const int p = min(sub + 8 * r, 80);   // window position (the last lanes repeat position 80)
                const int yy = min(max(cy + p / 9 - 4, 0), h - 1), xx = min(max(cx + p % 9 - 4, 0), w - 1);
                wv[r] = src >= 0 ? __ldg(sc + (size_t)yy * w + xx) : 0.f;   // a clamped position is inside the window too: same decision
            }
            bool bad = false;
#pragma unroll
            for (int r = 0; r < 11; ++r)
                if (wv[r] > cs) {
                    if (ROUND == 0) bad = true;
                    else {
                        const int p = min(sub + 8 * r, 80);
                        const int yy = min(max(cy + p / 9 - 4, 0), h - 1), xx = min(max(cx + p % 9 - 4, 0), w - 1);
                        if (!suppressed(yy, xx)) bad = true;
                    }
                }
            const unsigned bad_lanes = __ballot_sync(0xffffffffu, bad);
            const bool group_max = src >= 0 && ((bad_lanes >> (8 * grp)) & 0xFFu) == 0u;
            const unsigned max_groups = __ballot_sync(0xffffffffu, group_max && sub == 0);   // bit 8 g: group g found a maximum
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if (!((max_groups >> (8 * g)) & 1u)) continue;
                const int my = __shfl_sync(0xffffffffu, cy, 8 * g), mx = __shfl_sync(0xffffffffu, cx, 8 * g);
                const int ms = __shfl_sync(0xffffffffu, src, 8 * g);
                if (lane == ms) is_max = true;
                if (ROUND < 2 && lane < 9) {   // rows my - 4 .. my + 4, columns mx - 4 .. mx + 4 (clipped to the image)
                    const int yy = my - 4 + lane;
                    if (yy >= 0 && yy < h) {
                        const int x_lo = max(mx - 4, 0), x_hi = min(mx + 4, w - 1);
                        const int w0 = x_lo >> 5, w1 = x_hi >> 5;
                        const unsigned m_lo = 0xffffffffu << (x_lo & 31), m_hi = 0xffffffffu >> (31 - (x_hi & 31));
                        if (w0 == w1) {
                            atomicOr(&wb[yy * wpr + w0], m_lo & m_hi);
                            if (ROUND == 0) atomicOr(&wa[yy * wpr + w0], m_lo & m_hi);
                        } else {
                            atomicOr(&wb[yy * wpr + w0], m_lo); atomicOr(&wb[yy * wpr + w1], m_hi);
                            if (ROUND == 0) { atomicOr(&wa[yy * wpr + w0], m_lo); atomicOr(&wa[yy * wpr + w1], m_hi); }
                        }
                    }
                }
            }
        }
        // a maximum is final whatever the later rounds find: it becomes a top-K candidate now
        const bool keep = is_max && s > threshold && py >= border && py < h - border && px >= border && px < w - border;
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        if (ballot) {
            if (keep) s_keys[warp][n_buf + __popc(ballot & ((1u << lane) - 1))] =
                          ((unsigned long long)(0xFFFFFFFFu - __float_as_uint(s)) << 32) | (unsigned)(py * w + px);
            n_buf += __popc(ballot);
            __syncwarp();
            if (n_buf > NR_BUF - 32) flush();
        }
    }
    if (n_buf) flush();
}

// after the three rounds: redo_flag[slot] = 1 (and the candidate counter restarts) when the list overflowed or fewer than
// K survivors lie above a level that was higher than the plain threshold
__global__ void nms_redo_prepare_lists(int n, int slot0, int k_cap, unsigned thr_bits, const unsigned* __restrict__ level,
                                       const int* __restrict__ list_count, int* __restrict__ cand_count, int* __restrict__ redo_flag,
                                       int* __restrict__ any_redo) {
    // one CTA: any_redo lets the CTAs of the redo launch leave after ONE load when (as always with real frames) nothing is flagged
    int mine = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int slot = slot0 + i;
        const bool redo = list_count[i] > GNB_NMS_LIST_CAP || (cand_count[slot] < k_cap && level[slot] > thr_bits + 1u);
        redo_flag[slot] = redo ? 1 : 0;
        if (redo) { cand_count[slot] = 0; mine = 1; }
    }
    mine = __syncthreads_or(mine);
    if (threadIdx.x == 0) *any_redo = mine;
}

// One CTA per image: exact top-K of the candidate keys (radix select on the 64-bit key, 8 bits per pass, most
// significant byte first; the pass loop stops as soon as the selected bin holds exactly the keys still needed, which with
// distinct scores is after the four score bytes), then a bitonic sort of the selected keys: up to 1024 keys live one per
// thread and every compare-exchange with a partner less than 32 lanes away is a warp shuffle (40 of the 55 steps), only
// the wider ones go through shared memory.
__global__ void __launch_bounds__(1024) topk_kernel(const unsigned long long* __restrict__ cand_keys,
                                                    const int* __restrict__ cand_count, int slot0, int w, int k_cap,
                                                    float* __restrict__ kp_xy, float* __restrict__ kp_score,
                                                    int* __restrict__ kp_count) {
    __shared__ unsigned long long sel[GNB_MAX_KP];
    __shared__ int hist[256];
    __shared__ unsigned long long s_prefix, s_mask;
    __shared__ int s_remaining, s_nsel, s_done;
    const int slot = slot0 + blockIdx.x;
    const int n_raw = cand_count[slot];
    if (n_raw > GNB_CAND_CAP) {  // overflow: candidates were dropped in atomic order -> refuse
        if (threadIdx.x == 0) kp_count[slot] = -1;
        return;
    }
    const int n = n_raw;
    const int lane = threadIdx.x & 31;
    const unsigned long long* keys = cand_keys + (size_t)slot * GNB_CAND_CAP;
    unsigned long long kth = ~0ull;
    if (n > k_cap) {
        if (threadIdx.x == 0) { s_prefix = 0; s_mask = 0; s_remaining = k_cap; s_done = 0; }
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
            if (threadIdx.x < 32) {
                // warp 0: lane l owns bins [8 l, 8 l + 8); find the bin where the running count reaches s_remaining
                int mine = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) mine += hist[lane * 8 + j];
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t2 = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t2;
                }
                const int before = incl - mine, rem = s_remaining;
                __syncwarp();   // every lane has read s_remaining before the crossing lane rewrites it
                const bool crossing = before < rem && incl >= rem;   // exactly one lane (the total is >= rem by construction)
                if (crossing) {
                    int cum = before, d = lane * 8;
                    for (; d < lane * 8 + 7; ++d) {
                        if (cum + hist[d] >= rem) break;
                        cum += hist[d];
                    }
                    s_remaining = rem - cum;
                    s_prefix = prefix | ((unsigned long long)d << (8 * pass));
                    s_mask = mask | (255ull << (8 * pass));
                    // every key of the selected bin is needed: the K-th key is the largest key with this prefix
                    if (hist[d] == rem - cum) s_done = 1;
                }
            }
            __syncthreads();
            if (s_done) break;
        }
        kth = s_done ? (s_prefix | ~s_mask) : s_prefix;
    }
    if (threadIdx.x == 0) s_nsel = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        const unsigned long long key = i < n ? keys[i] : ~0ull;
        const bool take = i < n && key <= kth;
        const unsigned ballot = __ballot_sync(0xffffffffu, take);
        if (ballot) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_nsel, __popc(ballot));
            base = __shfl_sync(0xffffffffu, base, 0) + __popc(ballot & ((1u << lane) - 1));
            if (take && base < GNB_MAX_KP) sel[base] = key;
        }
    }
    __syncthreads();
    const int nsel = min(s_nsel, k_cap);
    int npow = 1;
    while (npow < nsel) npow <<= 1;
    if (npow <= 1024) {
        // one key per thread (threads >= nsel hold the +inf sentinel)
        unsigned long long key = (int)threadIdx.x < nsel ? sel[threadIdx.x] : ~0ull;
        const int i = threadIdx.x;
        for (int size = 2; size <= npow; size <<= 1) {
            const bool up = (i & size) == 0;
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                unsigned long long other;
                if (stride >= 32) {
                    __syncthreads();
                    sel[i] = key;
                    __syncthreads();
                    other = sel[i ^ stride];
                } else {
                    other = __shfl_xor_sync(0xffffffffu, key, stride);
                }
                const bool lower = (i & stride) == 0;          // this thread keeps the smaller key of the pair when sorting up
                const bool keep_min = lower == up;
                key = keep_min ? (other < key ? other : key) : (other > key ? other : key);
            }
        }
        __syncthreads();
        sel[i] = key;
        __syncthreads();
    } else {
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
    static const int dense_nms = getenv("GNB_NMS_DENSE") ? atoi(getenv("GNB_NMS_DENSE")) : 0;   // A/B: the dense generic-radius kernel
    if (ctx->cfg.nms_radius == 4 && ctx->cfg.keypoint_threshold >= 0.f && !dense_nms) {
        // sparse, top-K-aware path (see above)
        const int k_cap = ctx->cfg.max_keypoints;
        const unsigned thr_bits = __float_as_uint_host(ctx->cfg.keypoint_threshold);
        // listed pixels per image the level aims at.  Measured with the shipped weights (oracle, 720p / 1024^2 frames): 16 K
        // listed pixels hold 1.8 K survivors at K = 1024 and 1.3 K at K = 2048 (8 K: 1.2 K / 0.9 K — too close); an image
        // that still ends below K is redone from the threshold, so this number only steers the work, never the result
        static const int level_mult = getenv("GNB_NMS_LEVEL_MULT") ? max(1, atoi(getenv("GNB_NMS_LEVEL_MULT"))) : 16;
        const unsigned target = (unsigned)(((long long)level_mult * k_cap + NS_HIST_ROW_STEP - 1) / NS_HIST_ROW_STEP);
        const int rows = ceil_div(h, NS_HIST_ROW_STEP);
        // few CTAs per image when there are many images: every CTA ends with a 2048-bucket merge into the global histogram
        dim3 hgrid(min(ceil_div(rows * w, 256 * 4), max(8, min(128, ceil_div(ctx->sm_count * 4, n)))), 1, n);
        GNB_KERNEL(ctx, "nms_hist_kernel", nms_hist_kernel<<<hgrid, 256, 0, ctx->stream>>>(score, h, w, thr_bits, slot0, ctx->nms_hist,
                                                                                           ctx->nms_hist + (size_t)ctx->kp_slots * NS_BUCKETS, target, ctx->nms_level, ctx->nms_flag,
                                                                                           ctx->cand_count, ctx->nms_list_count + slot0));
        const size_t smem = (size_t)NS_REG * NS_PITCH * sizeof(float) + 3 * NS_REG * NS_WORDS * sizeof(unsigned) +
                            (size_t)(NS_LIST_CAP + NS_SURV_CAP) * sizeof(unsigned short);
        GNB_CUDA(ctx, gnb_func_smem(ctx, nms_sparse_kernel, (int)smem));
        const int items = ceil_div(w, NS_TILE) * ceil_div(h, NS_TILE) * n;
        const int grid = min(items, ctx->sm_count * 3);
        static const int tile_nms = getenv("GNB_NMS_TILES") ? atoi(getenv("GNB_NMS_TILES")) : 0;   // A/B: the tile kernel for the first pass too
        if (!tile_nms && (size_t)h * ceil_div(w, 32) <= ctx->nms_sup_words) {
            // list-based passes over the whole image (see nms_compact_kernel)
            const int hw = h * w, wpr = ceil_div(w, 32);
            // per-call scratch lives at the call's first slot, so that two calls on different slot ranges can run concurrently
            unsigned* sup_a = ctx->nms_sup + (size_t)slot0 * 2 * ctx->nms_sup_words;
            unsigned* sup_b = sup_a + (size_t)n * h * wpr;
            uint2* lists = ctx->nms_list + (size_t)slot0 * GNB_NMS_LIST_CAP;
            int* list_count = ctx->nms_list_count + slot0;
            GNB_CUDA(ctx, cudaMemsetAsync(sup_a, 0, sizeof(unsigned) * 2 * n * h * wpr, ctx->stream));
            dim3 cgrid(ceil_div(hw, 256 * NL_PER_THREAD), n);
            if ((hw & 3) == 0)
                GNB_KERNEL(ctx, "nms_compact_kernel", nms_compact_kernel<true><<<cgrid, 256, 0, ctx->stream>>>(score, hw, slot0, ctx->nms_level, lists, list_count));
            else
                GNB_KERNEL(ctx, "nms_compact_kernel", nms_compact_kernel<false><<<cgrid, 256, 0, ctx->stream>>>(score, hw, slot0, ctx->nms_level, lists, list_count));
            dim3 rgrid(max(1, ceil_div(ctx->sm_count * 8, n)), n);
            const float thr = ctx->cfg.keypoint_threshold;
            const int border = ctx->cfg.border;
            GNB_KERNEL(ctx, "nms_round_kernel<0>", nms_round_kernel<0><<<rgrid, 256, 0, ctx->stream>>>(score, h, w, slot0, lists, list_count, sup_a, sup_b, wpr,
                                                                                                      thr, border, ctx->cand_keys, ctx->cand_count));
            GNB_KERNEL(ctx, "nms_round_kernel<1>", nms_round_kernel<1><<<rgrid, 256, 0, ctx->stream>>>(score, h, w, slot0, lists, list_count, sup_a, sup_b, wpr,
                                                                                                      thr, border, ctx->cand_keys, ctx->cand_count));
            GNB_KERNEL(ctx, "nms_round_kernel<2>", nms_round_kernel<2><<<rgrid, 256, 0, ctx->stream>>>(score, h, w, slot0, lists, list_count, sup_a, sup_b, wpr,
                                                                                                      thr, border, ctx->cand_keys, ctx->cand_count));
            int* any_redo = ctx->nms_list_count + ctx->kp_slots + slot0;
            GNB_KERNEL(ctx, "nms_redo_prepare", nms_redo_prepare_lists<<<1, 256, 0, ctx->stream>>>(n, slot0, k_cap, thr_bits, ctx->nms_level, list_count,
                                                                                           ctx->cand_count, ctx->nms_flag, any_redo));
            GNB_KERNEL(ctx, "nms_sparse_kernel(redo)", nms_sparse_kernel<<<grid, NS_THREADS, smem, ctx->stream>>>(
                score, h, w, n, thr, border, slot0, ctx->nms_level, ctx->nms_flag, 2, k_cap, thr_bits, ctx->cand_keys, ctx->cand_count, any_redo));
        } else {
            GNB_KERNEL(ctx, "nms_sparse_kernel", nms_sparse_kernel<<<grid, NS_THREADS, smem, ctx->stream>>>(
                score, h, w, n, ctx->cfg.keypoint_threshold, ctx->cfg.border, slot0, ctx->nms_level, nullptr, 0, k_cap, thr_bits, ctx->cand_keys, ctx->cand_count));
            // exact redo, from the plain threshold, of the images with fewer than K survivors above their level (rare); its
            // CTAs decide from the first pass's counts (copied aside: the redo resets and refills cand_count) and exit at once
            // when there is nothing to redo
            GNB_KERNEL(ctx, "nms_redo_prepare", nms_redo_prepare<<<ceil_div(n, 64), 64, 0, ctx->stream>>>(n, slot0, k_cap, thr_bits, ctx->nms_level, ctx->cand_count,
                                                                                                   ctx->nms_flag));
            GNB_KERNEL(ctx, "nms_sparse_kernel(redo)", nms_sparse_kernel<<<grid, NS_THREADS, smem, ctx->stream>>>(
                score, h, w, n, ctx->cfg.keypoint_threshold, ctx->cfg.border, slot0, ctx->nms_level, ctx->nms_flag, 1, k_cap, thr_bits, ctx->cand_keys, ctx->cand_count));
        }
    } else {
        GNB_CUDA(ctx, cudaMemsetAsync(ctx->cand_count + slot0, 0, sizeof(int) * n, ctx->stream));
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
