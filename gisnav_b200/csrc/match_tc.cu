// match_tc.cu — K4 on the 5th-generation tensor cores: the descriptor GEMM S = m_a m_b^T runs as
// tcgen05.mma (bf16 operands staged by TMA into 128B-swizzled shared memory, fp32 accumulators in
// TMEM) with the dual-softmax statistics fused into the TMEM epilogue, so S never touches HBM.
//
// One CTA owns 128 rows of one side (128 TMEM lanes) and streams the other side in 128-column
// tiles through a 2-stage TMA ring and a 2-stage TMEM accumulator ring:
//   warp 0   TMA producer (one elected lane)        warp 1   MMA issuer (one elected lane)
//   warp 2   TMEM allocator                         warps 4-7 epilogue: tcgen05.ld, one row per thread
// pass 0: running (max, sum exp) per row -> row_lse.   pass 1: per-row argmax of the assignment
// score (oracle/matcher_ref.py) -> best_val / best_idx.  The "column" statistics are the row
// statistics of the swapped product (blockIdx.z = side), as in the SIMT validation kernel.
#include "tc_common.cuh"

#include <math.h>

#define MT_BM 128          // rows per CTA
#define MT_BN 128          // columns per tile
#define MT_K 256           // descriptor dimension
#define MT_KC 64           // K elements per 128-byte swizzle chunk
#define MT_CHUNK_BYTES (128 * 128)                 // 128 rows x 128 B
#define MT_TILE_BYTES (MT_CHUNK_BYTES * (MT_K / MT_KC))  // 64 KB: one operand tile, full K
#define MT_STAGES 2
#define MT_TMEM_COLS 256   // 2 accumulator stages x 128 columns
#define MT_EPI_GROUPS 2     // epilogue warpgroups: group g consumes columns [64 g, 64 g + 64) of every tile
#define MT_THREADS (128 + 128 * MT_EPI_GROUPS)
#define MT_SMEM_BYTES (1024 + MT_TILE_BYTES * (1 + MT_STAGES) + 256 + 2 * 2 * MT_BN * 4 + 3 * 128 * 4)

gnb_encode_tiled_fn gnb_get_encode_tiled(gnb_ctx* ctx) {
    static gnb_encode_tiled_fn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
        GNB_SET_ERR(ctx, "cuTensorMapEncodeTiled not available from the driver");
        return nullptr;
    }
    fn = reinterpret_cast<gnb_encode_tiled_fn>(p);
    return fn;
}

int gnb_make_tmap_bf16(gnb_ctx* ctx, CUtensorMap* out, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box) {
    gnb_encode_tiled_fn fn = gnb_get_encode_tiled(ctx);
    if (!fn) return GNB_E_CUDA;
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        GNB_SET_ERR(ctx, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
        return GNB_E_CUDA;
    }
    return GNB_OK;
}

void gnb_tc_state_free(gnb_ctx* ctx) {
    if (ctx->col_pa) { cudaFree(ctx->col_pa); ctx->col_pa = nullptr; }
    if (ctx->col_pb) { cudaFree(ctx->col_pb); ctx->col_pb = nullptr; }
    if (ctx->col_qa) { cudaFree(ctx->col_qa); ctx->col_qa = nullptr; }
    if (ctx->col_qb) { cudaFree(ctx->col_qb); ctx->col_qb = nullptr; }
    if (ctx->mproj_x3) { cudaFree(ctx->mproj_x3); ctx->mproj_x3 = nullptr; }
    if (ctx->tc_state) { delete static_cast<TcState*>(ctx->tc_state); ctx->tc_state = nullptr; }
    if (ctx->tc_err_host) { cudaFreeHost(ctx->tc_err_host); ctx->tc_err_host = nullptr; ctx->tc_err_dev = nullptr; }
}

int* gnb_tc_err_dev(gnb_ctx* ctx) {
    if (!ctx->tc_err_host) {
        // portable: visible to every CUDA context of the process; one word per gnb_ctx so that two contexts
        // (two GPUs, two host threads) never see each other's watchdog
        if (cudaHostAlloc((void**)&ctx->tc_err_host, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return nullptr;
        *ctx->tc_err_host = 0;
        if (cudaHostGetDevicePointer((void**)&ctx->tc_err_dev, ctx->tc_err_host, 0) != cudaSuccess) return nullptr;
    }
    return ctx->tc_err_dev;
}
int gnb_tc_err_check(gnb_ctx* ctx) {
    if (ctx->tc_err_host && *ctx->tc_err_host) {
        GNB_SET_ERR(ctx, "tcgen05 pipeline wait timed out (code %d)", *ctx->tc_err_host);
        *ctx->tc_err_host = 0;
        return GNB_E_CUDA;
    }
    return GNB_OK;
}

// Row pass for the TwistNode matcher: brute-force L2 2-NN of every row of side A against side B (d^2 = |a|^2 + |b|^2 - 2 a.b),
// the two smallest distances per row (lowest index on ties).  (The LightGlue head uses match_pair_tc below.)
__global__ void __launch_bounds__(MT_THREADS, 1) match_knn_tc(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ mlogit,
                                                        const int* __restrict__ kp_count, int k_cap, int slot_a0, int stride_a, int slot_b0, int max_pairs,
                                                        float* __restrict__ row_lse, float* __restrict__ best_val,
                                                        int* __restrict__ best_idx, int* err) {
    constexpr int PASS = 2;
    const int pair = blockIdx.y, side = blockIdx.z;
    const int slot_a = slot_a0 + pair * stride_a, slot_b = slot_b0 + pair;
    const int slot_r = side == 0 ? slot_a : slot_b, slot_c = side == 0 ? slot_b : slot_a;
    const int rs_r = side == 0 ? pair : max_pairs + pair, rs_c = side == 0 ? max_pairs + pair : pair;
    const int nr = max(kp_count[slot_r], 0), nc = max(kp_count[slot_c], 0);
    const int r0 = blockIdx.x * MT_BM;
    if (r0 >= nr || nc == 0) return;  // uniform exit before any barrier / TMEM allocation
    const int n_tiles = (nc + MT_BN - 1) / MT_BN;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                                   // 64 KB
    uint8_t* sB = smem + MT_TILE_BYTES;                   // MT_STAGES x 64 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MT_TILE_BYTES * (1 + MT_STAGES));
    uint64_t* a_full = bars;            // 1
    uint64_t* b_full = bars + 1;        // [2]
    uint64_t* b_empty = bars + 3;       // [2]
    uint64_t* t_full = bars + 5;        // [2]
    uint64_t* t_empty = bars + 7;       // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
    float* s_cl = reinterpret_cast<float*>(smem + MT_TILE_BYTES * (1 + MT_STAGES) + 256);  // [2][128] column LSE
    float* s_lb = s_cl + 2 * MT_BN;                                                       // [2][128] column logit
    float* s_mrg = s_lb + 2 * MT_BN;                                                      // [3][128] partials of epilogue group 1

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmap);
        tc::mbar_init(a_full, 1);
        for (int s = 0; s < MT_STAGES; ++s) {
            tc::mbar_init(&b_full[s], 1);
            tc::mbar_init(&b_empty[s], 1);
            tc::mbar_init(&t_full[s], 1);
            tc::mbar_init(&t_empty[s], 4 * MT_EPI_GROUPS);  // one arrive per epilogue warp
        }
        tc::fence_barrier_init();
    }
    if (warp == 2) {
        tc::tmem_alloc(tmem_slot, MT_TMEM_COLS);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(a_full, MT_TILE_BYTES);
            for (int c = 0; c < MT_K / MT_KC; ++c) tc::tma_load_3d(sA + c * MT_CHUNK_BYTES, &tmap, a_full, c * MT_KC, r0, slot_r);
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j % MT_STAGES;
                const uint32_t ph = (j / MT_STAGES) & 1;
                if (j >= MT_STAGES && !tc::mbar_wait(&b_empty[s], ph ^ 1, err, 101)) break;
                tc::mbar_arrive_expect_tx(&b_full[s], MT_TILE_BYTES);
                for (int c = 0; c < MT_K / MT_KC; ++c)
                    tc::tma_load_3d(sB + s * MT_TILE_BYTES + c * MT_CHUNK_BYTES, &tmap, &b_full[s], c * MT_KC, j * MT_BN, slot_c);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: warp converged, one elected lane issues =====
        const uint32_t idesc = tc::make_idesc_bf16(MT_BM, MT_BN);
        bool ok = tc::mbar_wait(a_full, 0, err, 102);
        const uint64_t da0 = tc::make_smem_desc_sw128(tc::smem_u32(sA), 1024);
        const uint64_t db00 = tc::make_smem_desc_sw128(tc::smem_u32(sB), 1024);
        for (int j = 0; ok && j < n_tiles; ++j) {
            const int s = j % MT_STAGES;
            const uint32_t ph = (j / MT_STAGES) & 1;
            if (!tc::mbar_wait(&b_full[s], ph, err, 103)) break;
            if (j >= MT_STAGES && !tc::mbar_wait(&t_empty[s], ph ^ 1, err, 104)) break;
            tc::tc_fence_after();
            if (tc::elect_one()) {
                const uint32_t d_tmem = tmem_base + (uint32_t)(s * MT_BN);
                const uint64_t db0 = db00 + (uint64_t)((s * MT_TILE_BYTES) >> 4);
#pragma unroll
                for (int c = 0; c < MT_K / MT_KC; ++c) {
#pragma unroll
                    for (int k = 0; k < MT_KC / 16; ++k) {
                        const uint64_t off = (uint64_t)((c * MT_CHUNK_BYTES + k * 32) >> 4);
                        tc::umma_bf16(d_tmem, da0 + off, db0 + off, idesc, (c | k) ? 1u : 0u);
                    }
                }
                tc::umma_commit(&b_empty[s]);  // smem stage reusable once these MMAs have read it
                tc::umma_commit(&t_full[s]);   // accumulator ready for the epilogue
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ===== epilogue: thread <-> TMEM lane <-> row =====
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row_local = q * 32 + lane;
        const int row = r0 + row_local;
        const int et = threadIdx.x - 128;       // 0..255 among the epilogue threads
        const int grp = et >> 7;                // which half of each tile's columns this thread consumes
        float run_max = -INFINITY, run_sum = 0.f, bv = -INFINITY;
        int bi = -1;
        float rl = 0.f, la = 0.f;
        if (PASS == 1 && row < nr) { rl = row_lse[(size_t)rs_r * k_cap + row]; la = mlogit[(size_t)slot_r * k_cap + row]; }
        if (PASS == 2 && row < nr) la = mlogit[(size_t)slot_r * k_cap + row];   // |a|^2
        float b2v = INFINITY;   // PASS 2: second-smallest squared distance (bv = smallest, bi = its column)
        if (PASS == 2) bv = INFINITY;
        for (int j = 0; j < n_tiles; ++j) {
            const int s = j % MT_STAGES;
            const uint32_t ph = (j / MT_STAGES) & 1;
            const int c0 = j * MT_BN;
            if (PASS >= 1) {
                if (et < MT_BN) {
                    const int col = c0 + et;
                    s_cl[s * MT_BN + et] = (PASS == 1 && col < nc) ? row_lse[(size_t)rs_c * k_cap + col] : 0.f;
                    s_lb[s * MT_BN + et] = col < nc ? mlogit[(size_t)slot_c * k_cap + col] : 0.f;
                }
                tc::named_bar_sync(1, 128 * MT_EPI_GROUPS);
            }
            if (!tc::mbar_wait(&t_full[s], ph, err, 105)) break;
            tc::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * MT_BN);
#pragma unroll 1
            for (int cc = grp * (MT_BN / MT_EPI_GROUPS); cc < (grp + 1) * (MT_BN / MT_EPI_GROUPS); cc += 32) {
                uint32_t v[32];
                tc::tmem_ld32(taddr + cc, v);
                tc::tmem_ld_wait();
                if (PASS == 0) {
                    float m = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c0 + cc + i < nc) m = fmaxf(m, __uint_as_float(v[i]));
                    if (m > -INFINITY) {   // a chunk made only of masked columns leaves the running pair untouched
                        const float nm = fmaxf(run_max, m);
                        float e = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (c0 + cc + i < nc) e += expf(__uint_as_float(v[i]) - nm);
                        run_sum = (run_max > -INFINITY ? run_sum * expf(run_max - nm) : 0.f) + e;
                        run_max = nm;
                    }
                } else if (PASS == 2) {
                    // brute-force L2: d^2 = |a|^2 + |b|^2 - 2 a.b ; keep the two nearest (lowest index on ties)
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int col = c0 + cc + i;
                        if (col < nc) {
                            const float d2 = __fsub_rn(__fadd_rn(la, s_lb[s * MT_BN + cc + i]), __fmul_rn(2.0f, __uint_as_float(v[i])));
                            if (d2 < bv) { b2v = bv; bv = d2; bi = col; }
                            else if (d2 < b2v) b2v = d2;
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int col = c0 + cc + i;
                        if (col < nc) {
                            const float sv = __uint_as_float(v[i]);
                            const float t_row = __fsub_rn(sv, rl), t_col = __fsub_rn(sv, s_cl[s * MT_BN + cc + i]);
                            const float lb = s_lb[s * MT_BN + cc + i];
                            float sc;
                            if (side == 0) sc = __fadd_rn(__fadd_rn(__fadd_rn(t_row, t_col), la), lb);
                            else sc = __fadd_rn(__fadd_rn(__fadd_rn(t_col, t_row), lb), la);
                            if (sc > bv || bi < 0) { bv = sc; bi = col; }
                        }
                    }
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&t_empty[s]);
        }
        // merge the two column halves of every row: group 1 publishes, group 0 combines and stores
        if (grp == 1) {
            s_mrg[row_local] = PASS == 0 ? run_max : bv;
            s_mrg[128 + row_local] = PASS == 0 ? run_sum : b2v;
            s_mrg[256 + row_local] = __int_as_float(bi);
        }
        tc::named_bar_sync(2, 128 * MT_EPI_GROUPS);
        if (grp == 0 && row < nr) {
            const float o0 = s_mrg[row_local], o1 = s_mrg[128 + row_local];
            const int oi = __float_as_int(s_mrg[256 + row_local]);
            if (PASS == 0) {
                const float nm = fmaxf(run_max, o0);
                const float tot = (run_max > -INFINITY ? run_sum * expf(run_max - nm) : 0.f) + (o0 > -INFINITY ? o1 * expf(o0 - nm) : 0.f);
                row_lse[(size_t)rs_r * k_cap + row] = nm + logf(tot);
            } else if (PASS == 1) {
                if (oi >= 0 && (bi < 0 || o0 > bv || (o0 == bv && oi < bi))) { bv = o0; bi = oi; }
                best_val[(size_t)rs_r * k_cap + row] = bv; best_idx[(size_t)rs_r * k_cap + row] = bi;
            } else {
                // two smallest of {bv, b2v, o0, o1}; nearest keeps the lowest column index on ties
                float n1 = bv, n2 = b2v; int ni = bi;
                if (o0 < n1 || (o0 == n1 && oi >= 0 && oi < ni)) { n2 = fminf(n1, o1); n1 = o0; ni = oi; }
                else n2 = fminf(n2, o0);
                best_val[(size_t)rs_r * k_cap + row] = n1; best_idx[(size_t)rs_r * k_cap + row] = ni;
                row_lse[(size_t)rs_r * k_cap + row] = n2;
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, MT_TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// K4, one S per pair and pass.  A CTA owns 128 rows of side A of one pair (resident in shared memory) and streams side B
// in 128-column tiles, 64-element K chunks through a TMA ring.  Per tile it issues TWO accumulations from the same
// shared-memory operands: D1 = A_i B_j^T (TMEM lane = row of S) and D2 = B_j A_i^T (TMEM lane = column of S), so both
// the row statistics and the column statistics of the dual softmax come out of "one row per thread" epilogues: no
// transposed reduction, half the operand traffic of the row-pass kernel above (which streamed each side against the
// other twice), and the matcher was L2 -> SM operand-traffic bound.  Epilogue warpgroup 0 reads D1, warpgroup 1 reads
// D2.  Column results are partial (this CTA's 128 rows): they go to [pair][row block][column] arrays that the next
// stage reduces in row-block order (deterministic; ties resolve to the lowest index like numpy's argmax).
//   PASS 0: row LSE (complete) + per-column (max, sum exp) partials.
//   PASS 1: per-row argmax of the assignment score (complete) + per-column argmax partials.
//   X3   : fp32-faithful mode.  m = hi + lo as two bf16 terms ([hi: 256 | lo: 256] per keypoint), S keeps
//          hi hi^T + hi lo^T + lo hi^T: side A resident as 8 chunks, side B streamed as hi0, lo0, hi1, lo1, ...
#define MP_CHUNK (128 * 128)   // bytes: 128 keypoints x 64 K elements
// exp(x) = ex2(x log2 e) on the MUFU unit (relative error ~2^-22): the accurate expf costs ~25 instructions per score and
// made the statistics pass issue-bound
__device__ __forceinline__ float mp_exp(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}
template <int PASS, bool X3>
__global__ void __launch_bounds__(384, 1) match_pair_tc(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ mlogit,
                                                        const int* __restrict__ kp_count, int k_cap, int slot_a0, int stride_a, int slot_b0,
                                                        int max_pairs, int n_rb_cap, float* __restrict__ row_lse, float* __restrict__ best_val,
                                                        int* __restrict__ best_idx, float* __restrict__ col_pa, float* __restrict__ col_pb,
                                                        float* __restrict__ col_qa, float* __restrict__ col_qb, int* err) {
    // col_pa / col_pb: (max, sum exp) partials written by PASS 0 and read by every CTA of PASS 1; col_qa / col_qb: the
    // argmax partials PASS 1 writes (separate arrays: other CTAs of the same launch are still reading the pass-0 ones)
    constexpr int NA = X3 ? 8 : 4;     // resident A chunks (X3: hi0..3 then lo0..3)
    constexpr int NBQ = X3 ? 8 : 4;    // B chunks per column tile (X3: hi0, lo0, hi1, lo1, ...)
    constexpr int RING = X3 ? 5 : 8;
    const int pair = blockIdx.y, rb = blockIdx.x;
    const int slot_a = slot_a0 + pair * stride_a, slot_b = slot_b0 + pair;
    const int rs_a = pair, rs_b = max_pairs + pair;
    const int nr = max(kp_count[slot_a], 0), nc = max(kp_count[slot_b], 0);
    const int r0 = rb * 128;
    if (r0 >= nr || nc == 0) return;   // uniform exit before any barrier / TMEM allocation
    const int n_tiles = (nc + 127) / 128;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + NA * MP_CHUNK;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + RING * MP_CHUNK);
    uint64_t* a_full = bars;                 // 1
    uint64_t* b_full = bars + 1;             // [RING]
    uint64_t* b_empty = b_full + RING;       // [RING]
    uint64_t* t_full = b_empty + RING;       // [2]
    uint64_t* t_empty = t_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
    float2* s_c2 = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2][128] {column LSE, column logit} of the tile
    float2* s_r2 = s_c2 + 256;                                                         // [128] {row LSE, row logit} of this CTA's rows

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmap);
        tc::mbar_init(a_full, 1);
        for (int s = 0; s < RING; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&t_full[s], 1); tc::mbar_init(&t_empty[s], 8); }
        tc::fence_barrier_init();
    }
    if (warp == 2) { tc::tmem_alloc(tmem_slot, 512); tc::tmem_relinquish(); }
    if (PASS == 1 && threadIdx.x >= 128 && threadIdx.x < 256) {
        const int t = threadIdx.x - 128, row = r0 + t;
        s_r2[t] = row < nr ? make_float2(row_lse[(size_t)rs_a * k_cap + row], mlogit[(size_t)slot_a * k_cap + row]) : make_float2(0.f, 0.f);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(a_full, NA * MP_CHUNK);
            for (int c = 0; c < NA; ++c) tc::tma_load_3d(sA + c * MP_CHUNK, &tmap, a_full, c * 64, r0, slot_a);
            int i = 0;
            for (int j = 0; j < n_tiles; ++j) {
                for (int q = 0; q < NBQ; ++q, ++i) {
                    const int s = i % RING;
                    if (i >= RING && !tc::mbar_wait(&b_empty[s], ((i / RING) & 1) ^ 1, err, 111)) return;
                    const int kc = X3 ? ((q & 1) * 256 + (q >> 1) * 64) : q * 64;
                    tc::mbar_arrive_expect_tx(&b_full[s], MP_CHUNK);
                    tc::tma_load_3d(sB + s * MP_CHUNK, &tmap, &b_full[s], kc, j * 128, slot_b);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = tc::make_idesc_bf16(128, 128);
        bool ok = tc::mbar_wait(a_full, 0, err, 112);
        const uint64_t da0 = tc::make_smem_desc_sw128(tc::smem_u32(sA), 1024);
        const uint64_t db0 = tc::make_smem_desc_sw128(tc::smem_u32(sB), 1024);
        int i = 0;
        for (int j = 0; ok && j < n_tiles; ++j) {
            const int as = j & 1;
            if (j >= 2 && !tc::mbar_wait(&t_empty[as], ((j >> 1) & 1) ^ 1, err, 113)) break;
            const uint32_t d1 = tmem_base + (uint32_t)(as * 256), d2 = d1 + 128u;
#pragma unroll
            for (int q = 0; q < NBQ; ++q, ++i) {
                const int s = i % RING;
                if (!tc::mbar_wait(&b_full[s], (i / RING) & 1, err, 114)) { ok = false; break; }
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint64_t db = db0 + (uint64_t)((s * MP_CHUNK) >> 4);
                    // A chunks multiplied with this B chunk: plain = {q}; X3 hi chunk c = {hi c, lo c}; X3 lo chunk c = {hi c}
                    const int c = X3 ? (q >> 1) : q;
                    const int n_terms = (X3 && !(q & 1)) ? 2 : 1;
#pragma unroll
                    for (int term = 0; term < (X3 ? 2 : 1); ++term) {
                        if (term < n_terms) {
                            const uint64_t da = da0 + (uint64_t)(((c + term * 4) * MP_CHUNK) >> 4);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                tc::umma_bf16(d1, da + 2 * k, db + 2 * k, idesc, (q | term | k) ? 1u : 0u);   // rows of S
                                tc::umma_bf16(d2, db + 2 * k, da + 2 * k, idesc, (q | term | k) ? 1u : 0u);   // columns of S
                            }
                        }
                    }
                    tc::umma_commit(&b_empty[s]);
                    if (q == NBQ - 1) tc::umma_commit(&t_full[as]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: warpgroup 0 (warps 4-7) = rows of S from D1, warpgroup 1 (warps 8-11) = columns of S from D2 =====
        const int q4 = warp & 3, t = q4 * 32 + lane;
        const bool colgroup = warp >= 8;
        const int row = r0 + t;
        float run_max = -INFINITY, run_sum = 0.f;
        float bv4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        int bi4[4] = {-1, -1, -1, -1};
        auto merged_best = [&](float& bv, int& bi) {   // maximum of the four slots, ties to the lowest index
            bv = bv4[0]; bi = bi4[0];
#pragma unroll
            for (int q = 1; q < 4; ++q)
                if (bi4[q] >= 0 && (bi < 0 || bv4[q] > bv || (bv4[q] == bv && bi4[q] < bi))) { bv = bv4[q]; bi = bi4[q]; }
        };
        const float rl = (PASS == 1 && !colgroup) ? s_r2[t].x : 0.f, la = (PASS == 1 && !colgroup) ? s_r2[t].y : 0.f;
        float cl_next = 0.f, lb_next = 0.f;
        if (PASS == 1 && colgroup && t < nc) { cl_next = row_lse[(size_t)rs_b * k_cap + t]; lb_next = mlogit[(size_t)slot_b * k_cap + t]; }
        for (int j = 0; j < n_tiles; ++j) {
            const int as = j & 1, c0 = j * 128;
            const int col = c0 + t;     // colgroup: the column this thread owns in this tile
            float cl = 0.f, lb = 0.f;
            if (PASS == 1) {
                if (colgroup) {
                    cl = cl_next; lb = lb_next;
                    const int coln = col + 128;        // prefetch the next tile's column statistics
                    if (j + 1 < n_tiles && coln < nc) { cl_next = row_lse[(size_t)rs_b * k_cap + coln]; lb_next = mlogit[(size_t)slot_b * k_cap + coln]; }
                    s_c2[as * 128 + t] = make_float2(cl, lb);
#pragma unroll
                    for (int q = 0; q < 4; ++q) { bv4[q] = -INFINITY; bi4[q] = -1; }   // per-tile partial for this column
                }
                tc::named_bar_sync(1, 256);    // the row group reads s_cl / s_lb of this tile
            }
            if (!tc::mbar_wait(&t_full[as], (j >> 1) & 1, err, 115)) break;
            tc::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(as * 256 + (colgroup ? 128 : 0));
            float cmax = -INFINITY, csum = 0.f;   // colgroup, PASS 0: this tile's partial
#pragma unroll 1
            for (int cc = 0; cc < 128; cc += 32) {
                uint32_t v[32];
                tc::tmem_ld32(taddr + cc, v);
                tc::tmem_ld_wait();
                // "other" index of element i: rowgroup -> column c0 + cc + i (valid < nc); colgroup -> row r0 + cc + i (valid < nr)
                const int o0 = colgroup ? r0 + cc : c0 + cc, olim = colgroup ? nr : nc;
                if (PASS == 0) {
                    float m = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (o0 + i < olim) m = fmaxf(m, __uint_as_float(v[i]));
                    if (m > -INFINITY) {
                        float& rmax = colgroup ? cmax : run_max;
                        float& rsum = colgroup ? csum : run_sum;
                        const float nm = fmaxf(rmax, m);
                        float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;   // four independent partial sums
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            if (o0 + i < olim) e0 += mp_exp(__uint_as_float(v[i]) - nm);
                            if (o0 + i + 1 < olim) e1 += mp_exp(__uint_as_float(v[i + 1]) - nm);
                            if (o0 + i + 2 < olim) e2 += mp_exp(__uint_as_float(v[i + 2]) - nm);
                            if (o0 + i + 3 < olim) e3 += mp_exp(__uint_as_float(v[i + 3]) - nm);
                        }
                        rsum = (rmax > -INFINITY ? rsum * mp_exp(rmax - nm) : 0.f) + ((e0 + e1) + (e2 + e3));
                        rmax = nm;
                    }
                } else {
                    // score(i, j) = (((S - LSE_row_i) + (S - LSE_col_j)) + z_i) + z_j  (oracle/matcher_ref.py), same order in both
                    // groups.  All 32 scores first (independent), then four interleaved running maxima (element i -> slot
                    // i & 3) so the compare/select chain is 8 deep instead of 32; each slot sees ascending indices, so the
                    // strict > keeps the lowest index and the final merge breaks ties by index.
                    const float2* oth = colgroup ? s_r2 + cc : s_c2 + as * 128 + cc;   // {LSE, logit} of the other index
                    float sc[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float sv = __uint_as_float(v[i]);
                        const float2 ot = oth[i];
                        const float t_row = __fsub_rn(sv, colgroup ? ot.x : rl);
                        const float t_col = __fsub_rn(sv, colgroup ? cl : ot.x);
                        sc[i] = __fadd_rn(__fadd_rn(__fadd_rn(t_row, t_col), colgroup ? ot.y : la), colgroup ? lb : ot.y);
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (o0 + i < olim && (sc[i] > bv4[i & 3] || bi4[i & 3] < 0)) { bv4[i & 3] = sc[i]; bi4[i & 3] = o0 + i; }
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&t_empty[as]);
            if (colgroup && col < nc) {
                const size_t o = ((size_t)pair * n_rb_cap + rb) * k_cap + col;
                if (PASS == 0) { col_pa[o] = cmax; col_pb[o] = csum; }
                else { float bv; int bi; merged_best(bv, bi); col_qa[o] = bv; col_qb[o] = __int_as_float(bi); }
            }
        }
        if (!colgroup && row < nr) {
            if (PASS == 0) row_lse[(size_t)rs_a * k_cap + row] = run_max + logf(run_sum);
            else { float bv; int bi; merged_best(bv, bi); best_val[(size_t)rs_a * k_cap + row] = bv; best_idx[(size_t)rs_a * k_cap + row] = bi; }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}

// column LSE of S from the row-block partials of pass 0, combined in row-block order -> row_lse[max_pairs + pair][col]
__global__ void __launch_bounds__(256) match_collse_kernel(const float* __restrict__ col_pa, const float* __restrict__ col_pb,
                                                           const int* __restrict__ kp_count, int k_cap, int slot_a0, int stride_a, int slot_b0,
                                                           int max_pairs, int n_rb_cap, float* __restrict__ row_lse) {
    const int pair = blockIdx.y, col = blockIdx.x * blockDim.x + threadIdx.x;
    const int nr = max(kp_count[slot_a0 + pair * stride_a], 0), nc = max(kp_count[slot_b0 + pair], 0);
    if (col >= nc || nr == 0) return;
    const int n_rb_a = (nr + 127) / 128;
    float m = -INFINITY;
    for (int b2 = 0; b2 < n_rb_a; ++b2) m = fmaxf(m, col_pa[((size_t)pair * n_rb_cap + b2) * k_cap + col]);
    float tot = 0.f;
    for (int b2 = 0; b2 < n_rb_a; ++b2) {
        const size_t o = ((size_t)pair * n_rb_cap + b2) * k_cap + col;
        tot += col_pb[o] * mp_exp(col_pa[o] - m);
    }
    row_lse[(size_t)(max_pairs + pair) * k_cap + col] = m + logf(tot);
}

template <bool X3>
static constexpr int mp_smem_bytes() { return 1024 + ((X3 ? 8 : 4) + (X3 ? 5 : 8)) * MP_CHUNK + 256 + (256 + 128) * 8; }

// both passes for `pairs` pairs; column partials in ctx->col_pa / col_pb ([max_batch][n_rb_cap][K])
int gnb_match_tc_pairpass(gnb_ctx* ctx, int pairs, int slot_a0, int slot_b0, int stride_a) {
    const int k = ctx->cfg.max_keypoints;
    const int n_rb = ceil_div(k, 128);
    dim3 grid(n_rb, pairs);
    int* err = gnb_tc_err_dev(ctx);
    if (ctx->cfg.precision == 1) {
        const CUtensorMap& tm = tc_state(ctx)->match_map_x3;
        GNB_KERNEL(ctx, "match_pair_x3<0>", match_pair_tc<0, true><<<grid, 384, mp_smem_bytes<true>(), ctx->stream>>>(
            tm, ctx->mlogit, ctx->kp_count, k, slot_a0, stride_a, slot_b0, ctx->cfg.max_batch, n_rb, ctx->row_lse, ctx->best_val, ctx->best_idx, ctx->col_pa, ctx->col_pb, ctx->col_qa, ctx->col_qb, err));
        GNB_KERNEL(ctx, "match_collse_kernel", match_collse_kernel<<<dim3(ceil_div(k, 256), pairs), 256, 0, ctx->stream>>>(
            ctx->col_pa, ctx->col_pb, ctx->kp_count, k, slot_a0, stride_a, slot_b0, ctx->cfg.max_batch, n_rb, ctx->row_lse));
        GNB_KERNEL(ctx, "match_pair_x3<1>", match_pair_tc<1, true><<<grid, 384, mp_smem_bytes<true>(), ctx->stream>>>(
            tm, ctx->mlogit, ctx->kp_count, k, slot_a0, stride_a, slot_b0, ctx->cfg.max_batch, n_rb, ctx->row_lse, ctx->best_val, ctx->best_idx, ctx->col_pa, ctx->col_pb, ctx->col_qa, ctx->col_qb, err));
    } else {
        const CUtensorMap& tm = tc_state(ctx)->match_map;
        GNB_KERNEL(ctx, "match_pair_tc<0>", match_pair_tc<0, false><<<grid, 384, mp_smem_bytes<false>(), ctx->stream>>>(
            tm, ctx->mlogit, ctx->kp_count, k, slot_a0, stride_a, slot_b0, ctx->cfg.max_batch, n_rb, ctx->row_lse, ctx->best_val, ctx->best_idx, ctx->col_pa, ctx->col_pb, ctx->col_qa, ctx->col_qb, err));
        GNB_KERNEL(ctx, "match_collse_kernel", match_collse_kernel<<<dim3(ceil_div(k, 256), pairs), 256, 0, ctx->stream>>>(
            ctx->col_pa, ctx->col_pb, ctx->kp_count, k, slot_a0, stride_a, slot_b0, ctx->cfg.max_batch, n_rb, ctx->row_lse));
        GNB_KERNEL(ctx, "match_pair_tc<1>", match_pair_tc<1, false><<<grid, 384, mp_smem_bytes<false>(), ctx->stream>>>(
            tm, ctx->mlogit, ctx->kp_count, k, slot_a0, stride_a, slot_b0, ctx->cfg.max_batch, n_rb, ctx->row_lse, ctx->best_val, ctx->best_idx, ctx->col_pa, ctx->col_pb, ctx->col_qa, ctx->col_qb, err));
    }
    return GNB_OK;
}

int gnb_match_tc_init(gnb_ctx* ctx) {
    if (!gnb_tc_err_dev(ctx)) { GNB_SET_ERR(ctx, "cannot allocate the host-mapped error word"); return GNB_E_CUDA; }
    CUtensorMap* g_tmap_host = &tc_state(ctx)->match_map;
    const uint64_t k = (uint64_t)ctx->cfg.max_keypoints;
    const uint64_t dims[3] = {MT_K, k, (uint64_t)ctx->kp_slots};
    const uint64_t strides[2] = {MT_K * 2, k * MT_K * 2};
    const uint32_t box[3] = {MT_KC, 128, 1};
    int rc = gnb_make_tmap_bf16(ctx, g_tmap_host, ctx->mproj, 3, dims, strides, box);
    if (rc) return rc;
    GNB_CUDA(ctx, cudaFuncSetAttribute(match_knn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SMEM_BYTES));
    GNB_CUDA(ctx, cudaFuncSetAttribute(match_pair_tc<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mp_smem_bytes<false>()));
    GNB_CUDA(ctx, cudaFuncSetAttribute(match_pair_tc<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mp_smem_bytes<false>()));
    const size_t n_rb = (size_t)ceil_div((int)k, 128);
    GNB_CUDA(ctx, cudaMalloc(&ctx->col_pa, (size_t)ctx->cfg.max_batch * n_rb * k * sizeof(float)));
    GNB_CUDA(ctx, cudaMalloc(&ctx->col_pb, (size_t)ctx->cfg.max_batch * n_rb * k * sizeof(float)));
    GNB_CUDA(ctx, cudaMalloc(&ctx->col_qa, (size_t)ctx->cfg.max_batch * n_rb * k * sizeof(float)));
    GNB_CUDA(ctx, cudaMalloc(&ctx->col_qb, (size_t)ctx->cfg.max_batch * n_rb * k * sizeof(float)));
    if (ctx->cfg.precision == 1) {
        // fp32-faithful mode: projected descriptors as two bf16 terms per value, [slot][K][hi: 256 | lo: 256]
        GNB_CUDA(ctx, cudaMalloc(&ctx->mproj_x3, (size_t)ctx->kp_slots * k * 512 * sizeof(bf16)));
        GNB_CUDA(ctx, cudaMemset(ctx->mproj_x3, 0, (size_t)ctx->kp_slots * k * 512 * sizeof(bf16)));
        const uint64_t xd[3] = {512, k, (uint64_t)ctx->kp_slots};
        const uint64_t xs[2] = {1024, k * 1024};
        if ((rc = gnb_make_tmap_bf16(ctx, &tc_state(ctx)->match_map_x3, ctx->mproj_x3, 3, xd, xs, box))) return rc;
        GNB_CUDA(ctx, cudaFuncSetAttribute(match_pair_tc<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mp_smem_bytes<true>()));
        GNB_CUDA(ctx, cudaFuncSetAttribute(match_pair_tc<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mp_smem_bytes<true>()));
    }
    return GNB_OK;
}

// brute-force 2-NN row pass (gnb_knn_ratio)
int gnb_match_tc_rowpass(gnb_ctx* ctx, int pairs, int slot_a0, int slot_b0, int stride_a, int pass) {
    if (pass != 2) { GNB_SET_ERR(ctx, "gnb_match_tc_rowpass: only the kNN pass is a row pass"); return GNB_E_INVALID; }
    const int k = ctx->cfg.max_keypoints;
    CUtensorMap* g_tmap_host = &tc_state(ctx)->match_map;
    dim3 grid(ceil_div(k, MT_BM), pairs, 1);
    GNB_KERNEL(ctx, "match_knn_tc", match_knn_tc<<<grid, MT_THREADS, MT_SMEM_BYTES, ctx->stream>>>(
        *g_tmap_host, ctx->mlogit, ctx->kp_count, k, slot_a0, stride_a, slot_b0, ctx->cfg.max_batch, ctx->row_lse, ctx->best_val, ctx->best_idx, gnb_tc_err_dev(ctx)));
    return GNB_OK;
}
