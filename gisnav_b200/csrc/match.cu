// match.cu — K4: LightGlue assignment head (oracle/matcher_ref.py): projection, dual log-softmax,
// mutual argmax, threshold.  Replaces self._matcher(...) at
// ros/gisnav/gisnav/core/pose_node.py:285-287 and the match gather at :296-297.
//
//   m = (W d + b) / 256^(1/4)  (bf16)      z = w_m . d + b_m
//   S = m_a m_b^T                           score_ij = (S_ij - LSE_j' S_ij') + (S_ij - LSE_i' S_i'j)
//                                                      + logsigmoid(z_a,i) + logsigmoid(z_b,j)
//
// S is never written to HBM: one pass computes the row log-sum-exps of S and of S^T (the "column"
// statistics are the row statistics of the swapped product), a second pass recomputes S tile by
// tile and keeps the per-row argmax.  This file has the orchestration plus the SIMT validation
// tile kernels (cfg.match_impl = 1); match_tc.cu has the tcgen05 tile kernels (product path).
#include "common.cuh"

#include <math.h>

// matcher head weights, repacked on the device from the blob: W bf16 [out][in] and its transpose [in][out], w_m bf16
__global__ void repack_match_kernel(const float* __restrict__ w, const float* __restrict__ mw, bf16* __restrict__ out_w, bf16* __restrict__ out_wt,
                                    bf16* __restrict__ out_mw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 256 * 256) return;
    const int o = i >> 8, in = i & 255;
    const bf16 v = __float2bfloat16_rn(w[i]);
    out_w[i] = v;
    out_wt[in * 256 + o] = v;
    if (i < 256) out_mw[i] = __float2bfloat16_rn(mw[i]);
}

int gnb_match_init(gnb_ctx* ctx, const float* proj_w, const float* proj_b, const float* m_w, const float* m_b) {
    GNB_CUDA(ctx, cudaMalloc(&ctx->match_w, 256 * 256 * 2));
    GNB_CUDA(ctx, cudaMalloc(&ctx->match_wt, 256 * 256 * 2));
    GNB_CUDA(ctx, cudaMalloc(&ctx->match_b, 256 * 4));
    GNB_CUDA(ctx, cudaMalloc(&ctx->match_mw, 256 * 2));
    GNB_KERNEL(ctx, "repack_match_kernel", repack_match_kernel<<<256, 256, 0, ctx->stream>>>(proj_w, m_w, ctx->match_w, ctx->match_wt, ctx->match_mw));
    GNB_CUDA(ctx, cudaMemcpyAsync(ctx->match_b, proj_b, 256 * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    GNB_CUDA(ctx, cudaMemcpyAsync(&ctx->match_mb, m_b, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->cfg.precision == 1) {   // fp32-faithful mode: the head runs in fp32 on the CUDA cores
        GNB_CUDA(ctx, cudaMalloc(&ctx->match_w_f32, 256 * 256 * 4));
        GNB_CUDA(ctx, cudaMalloc(&ctx->match_mw_f32, 256 * 4));
        GNB_CUDA(ctx, cudaMemcpyAsync(ctx->match_w_f32, proj_w, 256 * 256 * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        GNB_CUDA(ctx, cudaMemcpyAsync(ctx->match_mw_f32, m_w, 256 * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    GNB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // match_mb is a host field
    return GNB_OK;
}

void gnb_match_free(gnb_ctx* ctx) {
    if (ctx->match_w) cudaFree(ctx->match_w);
    if (ctx->match_b) cudaFree(ctx->match_b);
    if (ctx->match_mw) cudaFree(ctx->match_mw);
    if (ctx->match_wt) { cudaFree(ctx->match_wt); ctx->match_wt = nullptr; }
    if (ctx->match_w_f32) { cudaFree(ctx->match_w_f32); ctx->match_w_f32 = nullptr; }
    if (ctx->match_mw_f32) { cudaFree(ctx->match_mw_f32); ctx->match_mw_f32 = nullptr; }
    ctx->match_w = nullptr; ctx->match_b = nullptr; ctx->match_mw = nullptr;
}

// ------------------------------------------------------------------------------------------------
// projection: 8 keypoints per CTA, thread o owns output channel o.
__global__ void __launch_bounds__(256) project_kernel(const float* __restrict__ desc, const int* __restrict__ kp_count,
                                                      int slot0, int k_cap, const bf16* __restrict__ wt,
                                                      const float* __restrict__ bias, const bf16* __restrict__ mw,
                                                      float mb, bf16* __restrict__ mproj, float* __restrict__ mlogit) {
    __shared__ float d[8][256];
    const int slot = slot0 + blockIdx.y;
    const int n = max(kp_count[slot], 0);
    const int kp0 = blockIdx.x * 8;
    if (kp0 >= n) return;
    const int o = threadIdx.x;
    for (int r = 0; r < 8; ++r) {
        float v = 0.f;
        if (kp0 + r < n) v = __bfloat162float(__float2bfloat16_rn(desc[((size_t)slot * k_cap + kp0 + r) * 256 + o]));
        d[r][o] = v;
    }
    __syncthreads();
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    for (int i = 0; i < 256; ++i) {
        const float w = __bfloat162float(wt[i * 256 + o]);
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = fmaf(d[r][i], w, acc[r]);
    }
    const float b = bias[o];
#pragma unroll
    for (int r = 0; r < 8; ++r)
        if (kp0 + r < n)
            mproj[((size_t)slot * k_cap + kp0 + r) * 256 + o] = __float2bfloat16_rn(__fmul_rn(__fadd_rn(acc[r], b), 0.25f));
    // matchability: warp r reduces keypoint r
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (kp0 + wid < n) {
        float z = 0.f;
        for (int i = lane; i < 256; i += 32) z = fmaf(d[wid][i], __bfloat162float(mw[i]), z);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) z += __shfl_xor_sync(0xffffffffu, z, s);
        z += mb;
        if (lane == 0) mlogit[(size_t)slot * k_cap + kp0 + wid] = fminf(z, 0.f) - log1pf(expf(-fabsf(z)));
    }
}

int gnb_project_tc(gnb_ctx* ctx, int slot0, int n_slots);
int gnb_project_f32(gnb_ctx* ctx, int slot0, int n_slots);   // conv_x3.cu

int gnb_match_project(gnb_ctx* ctx, int slot0, int n_slots) {
    if (ctx->cfg.precision == 1) return gnb_project_f32(ctx, slot0, n_slots);
    if (ctx->cfg.match_impl == 0) return gnb_project_tc(ctx, slot0, n_slots);
    const int k = ctx->cfg.max_keypoints;
    dim3 grid(ceil_div(k, 8), n_slots);
    GNB_KERNEL(ctx, "project_kernel", project_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->desc_f32, ctx->kp_count, slot0, k, ctx->match_wt, ctx->match_b,
                                                  ctx->match_mw, ctx->match_mb, ctx->mproj, ctx->mlogit));
    return GNB_OK;
}

// ------------------------------------------------------------------------------------------------
// SIMT tile pass.  CTA = 64 rows of side R against all columns of side C, 64 columns at a time;
// thread (ty,tx) owns a 4x4 micro-tile.  PASS 0: row log-sum-exp.  PASS 1: row argmax of the full
// assignment score.
__device__ __forceinline__ float mt_load(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float mt_load(const float* p) { return *p; }

// T = bf16 (validation of the tcgen05 kernels) or float (the fp32-faithful mode's matcher: cfg.precision = 1)
template <int PASS, typename T>
__global__ void __launch_bounds__(256) match_rows_simt(const T* __restrict__ mproj, const float* __restrict__ mlogit,
                                                       const int* __restrict__ kp_count, int k_cap, int slot_a0,
                                                       int stride_a, int slot_b0, int max_pairs, float* __restrict__ row_lse,
                                                       float* __restrict__ best_val, int* __restrict__ best_idx) {
    extern __shared__ float smf[];
    float* As = smf;              // [256][64]  (k-major, transposed)
    float* Bs = smf + 256 * 64;   // [256][64]
    const int pair = blockIdx.y, side = blockIdx.z;
    // data slots (a side may be shared by all pairs: stride_a = 0) and per-(pair, side) result rows
    const int slot_a = slot_a0 + pair * stride_a, slot_b = slot_b0 + pair;
    const int slot_r = side == 0 ? slot_a : slot_b, slot_c = side == 0 ? slot_b : slot_a;
    const int rs_r = side == 0 ? pair : max_pairs + pair, rs_c = side == 0 ? max_pairs + pair : pair;
    const int nr = max(kp_count[slot_r], 0), nc = max(kp_count[slot_c], 0);
    const int r0 = blockIdx.x * 64;
    if (r0 >= nr) return;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const T* A = mproj + (size_t)slot_r * k_cap * 256;
    const T* B = mproj + (size_t)slot_c * k_cap * 256;
    for (int i = tid; i < 64 * 256; i += 256) {
        const int r = i >> 8, k = i & 255;
        As[k * 64 + r] = (r0 + r < nr) ? mt_load(&A[(size_t)(r0 + r) * 256 + k]) : 0.f;
    }
    float run_max[4], run_sum[4], bv[4], rl[4], la[4];
    int bi[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        run_max[r] = -INFINITY; run_sum[r] = 0.f; bv[r] = -INFINITY; bi[r] = -1;
        const int row = r0 + ty * 4 + r;
        rl[r] = (PASS == 1 && row < nr) ? row_lse[(size_t)rs_r * k_cap + row] : 0.f;
        la[r] = (PASS == 1 && row < nr) ? mlogit[(size_t)slot_r * k_cap + row] : 0.f;
    }
    for (int c0 = 0; c0 < nc; c0 += 64) {
        __syncthreads();
        for (int i = tid; i < 64 * 256; i += 256) {
            const int c = i >> 8, k = i & 255;
            Bs[k * 64 + c] = (c0 + c < nc) ? mt_load(&B[(size_t)(c0 + c) * 256 + k]) : 0.f;
        }
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
#pragma unroll 4
        for (int k = 0; k < 256; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(As + k * 64 + ty * 4);
            const float4 b = *reinterpret_cast<const float4*>(Bs + k * 64 + tx * 4);
            const float av[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bw[c], acc[r][c]);
        }
        if (PASS == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float m = -INFINITY;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c0 + tx * 4 + c < nc) m = fmaxf(m, acc[r][c]);
#pragma unroll
                for (int s = 8; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
                const float nm = fmaxf(run_max[r], m);
                float e = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c0 + tx * 4 + c < nc) e += expf(acc[r][c] - nm);
#pragma unroll
                for (int s = 8; s > 0; s >>= 1) e += __shfl_xor_sync(0xffffffffu, e, s);
                run_sum[r] = run_sum[r] * expf(run_max[r] - nm) + e;
                run_max[r] = nm;
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float v = -INFINITY;
                int idx = -1;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int col = c0 + tx * 4 + c;
                    if (col < nc) {
                        const float s = acc[r][c];
                        const float cl = row_lse[(size_t)rs_c * k_cap + col];
                        const float lb = mlogit[(size_t)slot_c * k_cap + col];
                        // side 0: rows are a (softmax over dim 1 uses rl), side 1: rows are b
                        const float t_row = __fsub_rn(s, rl[r]), t_col = __fsub_rn(s, cl);
                        float sc;
                        if (side == 0) sc = __fadd_rn(__fadd_rn(__fadd_rn(t_row, t_col), la[r]), lb);
                        else sc = __fadd_rn(__fadd_rn(__fadd_rn(t_col, t_row), lb), la[r]);
                        if (sc > v) { v = sc; idx = col; }
                    }
                }
#pragma unroll
                for (int s = 8; s > 0; s >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, v, s);
                    const int oi = __shfl_xor_sync(0xffffffffu, idx, s);
                    if (oi >= 0 && (ov > v || (ov == v && oi < idx) || idx < 0)) { v = ov; idx = oi; }
                }
                if (idx >= 0 && (v > bv[r] || bi[r] < 0)) { bv[r] = v; bi[r] = idx; }
            }
        }
    }
    if (tx == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = r0 + ty * 4 + r;
            if (row < nr) {
                if (PASS == 0) row_lse[(size_t)rs_r * k_cap + row] = run_max[r] + logf(run_sum[r]);
                else { best_val[(size_t)rs_r * k_cap + row] = bv[r]; best_idx[(size_t)rs_r * k_cap + row] = bi[r]; }
            }
        }
    }
}

// mutual check + threshold + ordered compaction; one CTA per pair.
// col_pa / col_pb (optional): per-(row block, column) argmax partials of match_pair_tc<1>; the best row of a column is
// their maximum, ties to the lowest row block (= lowest row index, as within a block)
__global__ void __launch_bounds__(1024) mutual_kernel(const float* __restrict__ best_val, const int* __restrict__ best_idx,
                                                      const int* __restrict__ kp_count, const float* __restrict__ kp_xy,
                                                      int k_cap, int slot_a0, int stride_a, int slot_b0, int max_pairs, float thr,
                                                      int* __restrict__ match_idx, float* __restrict__ match_score,
                                                      int* __restrict__ match_count, float* __restrict__ mkp_qry,
                                                      float* __restrict__ mkp_ref, const float* __restrict__ col_pa,
                                                      const float* __restrict__ col_pb, int n_rb_cap) {
    __shared__ int warp_sums[32];
    __shared__ int s_base;
    const int pair = blockIdx.x, sa = slot_a0 + pair * stride_a, sb = slot_b0 + pair;
    const int ra = pair, rb = max_pairs + pair;  // result rows of the two sides
    const int na = max(kp_count[sa], 0), nb = max(kp_count[sb], 0);
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < na; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        bool ok = false;
        int j = -1;
        float ms = 0.f;
        if (i < na && nb > 0) {
            j = best_idx[(size_t)ra * k_cap + i];
            int back = -1;
            if (j >= 0) {
                if (col_pa) {
                    float bvv = -INFINITY;
                    const int n_rb_a = (na + 127) / 128;
                    for (int b2 = 0; b2 < n_rb_a; ++b2) {
                        const size_t o = ((size_t)pair * n_rb_cap + b2) * k_cap + j;
                        const float v = col_pa[o];
                        if (v > bvv || back < 0) { bvv = v; back = __float_as_int(col_pb[o]); }
                    }
                } else {
                    back = best_idx[(size_t)rb * k_cap + j];
                }
            }
            if (j >= 0 && back == i) {
                ms = expf(best_val[(size_t)ra * k_cap + i]);
                ok = ms > thr;
            }
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, ok);
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) warp_sums[wid] = __popc(ballot);
        __syncthreads();
        int offset = s_base;
        for (int q = 0; q < wid; ++q) offset += warp_sums[q];
        const int pos = offset + __popc(ballot & ((1u << lane) - 1));
        if (ok) {
            match_idx[((size_t)pair * k_cap + pos) * 2 + 0] = i;
            match_idx[((size_t)pair * k_cap + pos) * 2 + 1] = j;
            match_score[(size_t)pair * k_cap + pos] = ms;
            mkp_qry[((size_t)pair * k_cap + pos) * 2 + 0] = kp_xy[((size_t)sa * k_cap + i) * 2 + 0];
            mkp_qry[((size_t)pair * k_cap + pos) * 2 + 1] = kp_xy[((size_t)sa * k_cap + i) * 2 + 1];
            mkp_ref[((size_t)pair * k_cap + pos) * 2 + 0] = kp_xy[((size_t)sb * k_cap + j) * 2 + 0];
            mkp_ref[((size_t)pair * k_cap + pos) * 2 + 1] = kp_xy[((size_t)sb * k_cap + j) * 2 + 1];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int q = 0; q < 32; ++q) tot += warp_sums[q];
            s_base += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) match_count[pair] = s_base;
}

int gnb_match_pairs(gnb_ctx* ctx, int pairs, int slot_a0, int slot_b0, int stride_a) {
    const int mp = ctx->cfg.max_batch;
    const int k = ctx->cfg.max_keypoints;
    const size_t smem = 2 * 256 * 64 * sizeof(float);
    dim3 grid(ceil_div(k, 64), pairs, 2);
    bool partials = false;
    if (ctx->cfg.match_impl == 0) {
        // tcgen05: one S per pair and pass (bf16 operands, or split-bf16 x3 in the fp32-faithful mode)
        int rc;
        if ((rc = gnb_match_tc_pairpass(ctx, pairs, slot_a0, slot_b0, stride_a))) return rc;
        partials = true;
    } else if (ctx->cfg.precision == 1) {
        GNB_CUDA(ctx, gnb_func_smem(ctx, match_rows_simt<0, float>, (int)smem));
        GNB_CUDA(ctx, gnb_func_smem(ctx, match_rows_simt<1, float>, (int)smem));
        GNB_KERNEL(ctx, "match_rows_f32<0>", match_rows_simt<0, float><<<grid, 256, smem, ctx->stream>>>(ctx->mproj_f32, ctx->mlogit, ctx->kp_count, k, slot_a0, stride_a, slot_b0, mp,
                                                             ctx->row_lse, ctx->best_val, ctx->best_idx));
        GNB_KERNEL(ctx, "match_rows_f32<1>", match_rows_simt<1, float><<<grid, 256, smem, ctx->stream>>>(ctx->mproj_f32, ctx->mlogit, ctx->kp_count, k, slot_a0, stride_a, slot_b0, mp,
                                                             ctx->row_lse, ctx->best_val, ctx->best_idx));
    } else {
        GNB_CUDA(ctx, gnb_func_smem(ctx, match_rows_simt<0, bf16>, (int)smem));
        GNB_CUDA(ctx, gnb_func_smem(ctx, match_rows_simt<1, bf16>, (int)smem));
        GNB_KERNEL(ctx, "match_rows_simt<0>", match_rows_simt<0, bf16><<<grid, 256, smem, ctx->stream>>>(ctx->mproj, ctx->mlogit, ctx->kp_count, k, slot_a0, stride_a, slot_b0, mp,
                                                             ctx->row_lse, ctx->best_val, ctx->best_idx));
        GNB_KERNEL(ctx, "match_rows_simt<1>", match_rows_simt<1, bf16><<<grid, 256, smem, ctx->stream>>>(ctx->mproj, ctx->mlogit, ctx->kp_count, k, slot_a0, stride_a, slot_b0, mp,
                                                             ctx->row_lse, ctx->best_val, ctx->best_idx));
    }
    GNB_KERNEL(ctx, "mutual_kernel", mutual_kernel<<<pairs, 1024, 0, ctx->stream>>>(ctx->best_val, ctx->best_idx, ctx->kp_count, ctx->kp_xy, k, slot_a0,
                                                   stride_a, slot_b0, mp, ctx->cfg.match_threshold, ctx->match_idx, ctx->match_score,
                                                   ctx->match_count, ctx->mkp_qry, ctx->mkp_ref, partials ? ctx->col_qa : nullptr,
                                                   partials ? ctx->col_qb : nullptr, ceil_div(k, 128)));
    return GNB_OK;
}

// ------------------------------------------------------------------------------------------------
// TwistNode's visual-odometry matcher (ros/gisnav/gisnav/core/twist_node.py:95,248,263-267):
// cv2.BFMatcher().knnMatch(desc_qry, desc_ref, k=2) + Lowe ratio test m.distance < 0.7 n.distance.
// The L2 distances come from the same tcgen05 descriptor GEMM as K4 (d^2 = |a|^2 + |b|^2 - 2 a.b).
// SIFT descriptors are integers 0..255 stored as float: exactly representable in bf16 and every
// partial sum is < 2^24, so d^2 is exact and the result is bit-identical to OpenCV's.
__global__ void __launch_bounds__(256) knn_prepare_kernel(const float* __restrict__ desc, int n, int dim, int slot, int k_cap,
                                                          bf16* __restrict__ out, float* __restrict__ norm2) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    float ss = 0.f;
    for (int c = lane; c < 256; c += 32) {
        float v = c < dim ? desc[(size_t)row * dim + c] : 0.f;
        const bf16 q = __float2bfloat16_rn(v);
        out[((size_t)slot * k_cap + row) * 256 + c] = q;
        const float qf = __bfloat162float(q);
        ss = fmaf(qf, qf, ss);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
    if (lane == 0) norm2[(size_t)slot * k_cap + row] = ss;
}

__global__ void __launch_bounds__(1024) ratio_kernel(const float* __restrict__ d1sq, const float* __restrict__ d2sq,
                                                     const int* __restrict__ j1, int n, double ratio, int* __restrict__ match_idx,
                                                     float* __restrict__ match_dist, int* __restrict__ match_count) {
    __shared__ int warp_sums[32];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        bool ok = false;
        float dist1 = 0.f;
        if (i < n) {
            dist1 = sqrtf(fmaxf(d1sq[i], 0.f));
            const float dist2 = sqrtf(fmaxf(d2sq[i], 0.f));
            ok = (double)dist1 < ratio * (double)dist2;   // Python: m.distance < 0.7 * n.distance (float64 arithmetic)
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, ok);
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) warp_sums[wid] = __popc(ballot);
        __syncthreads();
        int offset = s_base;
        for (int q = 0; q < wid; ++q) offset += warp_sums[q];
        const int pos = offset + __popc(ballot & ((1u << lane) - 1));
        if (ok) { match_idx[2 * pos] = i; match_idx[2 * pos + 1] = j1[i]; match_dist[pos] = dist1; }
        __syncthreads();
        if (threadIdx.x == 0) { int tot = 0; for (int q = 0; q < 32; ++q) tot += warp_sums[q]; s_base += tot; }
        __syncthreads();
    }
    if (threadIdx.x == 0) match_count[0] = s_base;
}

int gnb_knn_ratio(gnb_ctx* ctx, const float* dq, int nq, const float* dr, int nr, int dim, double ratio) {
    const int k = ctx->cfg.max_keypoints, sb = ctx->cfg.max_batch;
    GNB_KERNEL(ctx, "knn_prepare_kernel", knn_prepare_kernel<<<ceil_div(nq, 8), 256, 0, ctx->stream>>>(dq, nq, dim, 0, k, ctx->mproj, ctx->mlogit));
    GNB_KERNEL(ctx, "knn_prepare_kernel", knn_prepare_kernel<<<ceil_div(nr, 8), 256, 0, ctx->stream>>>(dr, nr, dim, sb, k, ctx->mproj, ctx->mlogit));
    int rc;
    if ((rc = gnb_match_tc_rowpass(ctx, 1, 0, sb, 1, 2))) return rc;
    GNB_KERNEL(ctx, "ratio_kernel", ratio_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->best_val, ctx->row_lse, ctx->best_idx, nq, ratio,
                                                                              ctx->match_idx, ctx->match_score, ctx->match_count));
    return GNB_OK;
}
