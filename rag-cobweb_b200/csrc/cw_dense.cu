// cw_dense.cu -- batched dense predict (CobwebWrapper.cobweb_predict_indexed /
// cobweb_rank_scores, src/cobweb/CobwebWrapper.py:210-294): every query of a batch against
// every node, path product, top-k.
//
// Kernel 1  dense_score_kernel: node_scores[q,b] = -0.5*(sumlog[b] + sum_d (x_qd*r_bd + mb_bd)^2)
//           with r = 1/sqrt(var), mb = -mean*r, i.e. the reference's (x-mean)^2/var evaluated as
//           TWO FFMAs per (query, node, attribute) on the FP32 pipe.  (x-mean) has to be formed
//           per (q, n, d) triple, so this is not a tensor-core contraction; the GEMM form
//           sum x^2/var - 2 sum x*mean/var + sum mean^2/var cancels catastrophically exactly
//           where ranking matters (query close to a leaf) -- DESIGN.md "Dense predict".
//           128x128 output tile per CTA, 8x8 per thread, operands pre-tiled k-major in HBM so
//           each pipeline stage is three contiguous 8 KB blocks fetched with cp.async.bulk
//           (TMA bulk copy, mbarrier completion), 4 stages.
// Kernel 2  paths_topk_kernel: leaf score = sequential FMA of path_w * node score, root first
//           (bit-identical to torch.sparse.mm on the reference's side), fused with a per-chunk
//           top-k; kernel 3 merges the chunk candidates.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cobweb_b200.h"

void cw_set_error(const char *fmt, ...);
int cw_check_cuda(cudaError_t e, const char *what);

namespace cw {

constexpr int TQ = 128, TN = CW_TILE_N, TK = CW_TILE_K, STAGES = 4, SCORE_THREADS = 256;

// ------------------------------------------------------------------ PTX helpers (mbarrier + bulk copy)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct __align__(128) ScoreStage {
    float x[TK][TQ];
    float r[TK][TN];
    float mb[TK][TN];
};

// grid: (query tiles, node tiles) -- query tile fastest so that concurrently resident CTAs
// share one node tile through L2 and the node matrices stream from HBM exactly once.
__global__ void __launch_bounds__(SCORE_THREADS, 2)
dense_score_kernel(const float *__restrict__ XT, const float *__restrict__ R, const float *__restrict__ MB,
                   const float *__restrict__ sumlog, float *__restrict__ out, long long ld, long long nq,
                   int n_ktiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ScoreStage *st = reinterpret_cast<ScoreStage *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + STAGES * sizeof(ScoreStage));
    uint64_t *empty = full + STAGES;

    const int tid = threadIdx.x;
    const int qt = blockIdx.x, nt = blockIdx.y;
    const float *xsrc = XT + (size_t)qt * n_ktiles * (TK * TQ);
    const float *rsrc = R + (size_t)nt * n_ktiles * (TK * TN);
    const float *msrc = MB + (size_t)nt * n_ktiles * (TK * TN);
    constexpr uint32_t XB = TK * TQ * 4, NB = TK * TN * 4;

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], SCORE_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < STAGES && s < n_ktiles; s++) {
            mbar_arrive_expect_tx(&full[s], XB + 2 * NB);
            bulk_g2s(st[s].x, xsrc + (size_t)s * (TK * TQ), XB, &full[s]);
            bulk_g2s(st[s].r, rsrc + (size_t)s * (TK * TN), NB, &full[s]);
            bulk_g2s(st[s].mb, msrc + (size_t)s * (TK * TN), NB, &full[s]);
        }
    }

    const int tx = tid & 15, ty = tid >> 4;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.0f;

    for (int kt = 0; kt < n_ktiles; kt++) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        mbar_wait(&full[s], ph);
        const ScoreStage &S = st[s];
#pragma unroll
        for (int kk = 0; kk < TK; kk++) {
            const float4 xa = *reinterpret_cast<const float4 *>(&S.x[kk][ty * 4]);
            const float4 xb = *reinterpret_cast<const float4 *>(&S.x[kk][64 + ty * 4]);
            const float4 ra = *reinterpret_cast<const float4 *>(&S.r[kk][tx * 4]);
            const float4 rb = *reinterpret_cast<const float4 *>(&S.r[kk][64 + tx * 4]);
            const float4 ma = *reinterpret_cast<const float4 *>(&S.mb[kk][tx * 4]);
            const float4 mb = *reinterpret_cast<const float4 *>(&S.mb[kk][64 + tx * 4]);
            const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            const float rv[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
            const float mv[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
#pragma unroll
            for (int i = 0; i < 8; i++) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float u = fmaf(xv[i], rv[j], mv[j]);
                    acc[i][j] = fmaf(u, u, acc[i][j]);
                }
            }
        }
        mbar_arrive(&empty[s]);
        if (tid == 0 && kt + STAGES < n_ktiles) {
            mbar_wait(&empty[s], ph);  // every thread is done reading this stage
            const int k2 = kt + STAGES;
            mbar_arrive_expect_tx(&full[s], XB + 2 * NB);
            bulk_g2s(st[s].x, xsrc + (size_t)k2 * (TK * TQ), XB, &full[s]);
            bulk_g2s(st[s].r, rsrc + (size_t)k2 * (TK * TN), NB, &full[s]);
            bulk_g2s(st[s].mb, msrc + (size_t)k2 * (TK * TN), NB, &full[s]);
        }
    }

    // epilogue: -0.5 * (sumlog + quad)   (CobwebWrapper.py:232-236)
    const int b0 = nt * TN + tx * 4;
    const float4 sla = *reinterpret_cast<const float4 *>(sumlog + b0);
    const float4 slb = *reinterpret_cast<const float4 *>(sumlog + b0 + 64);
    const float sl[8] = {sla.x, sla.y, sla.z, sla.w, slb.x, slb.y, slb.z, slb.w};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const long long q = (long long)qt * TQ + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (q < nq) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; j++) o[j] = -0.5f * (sl[j] + acc[i][j]);
            float *row = out + q * ld + b0;
            *reinterpret_cast<float4 *>(row) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4 *>(row + 64) = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
}

// Q [nq, D] row-major -> XT [q tile][k tile][TK][TQ], zero padded (same tiling as the index)
__global__ void __launch_bounds__(256)
tile_queries_kernel(const float *__restrict__ Q, long long nq, int D, int n_ktiles, float *XT) {
    __shared__ float t[TK][TQ + 1];
    const int qt = blockIdx.x, kt = blockIdx.y, tid = threadIdx.x;
    const int ql = tid >> 1, half = tid & 1;
    const long long q = (long long)qt * TQ + ql;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const int kk = half * 8 + e, d = kt * TK + kk;
        t[kk][ql] = (q < nq && d < D) ? Q[q * D + d] : 0.0f;
    }
    __syncthreads();
    const size_t tile = ((size_t)qt * n_ktiles + kt) * (TK * TQ);
    for (int i = tid; i < TK * TQ; i += 256) XT[tile + i] = t[i / TQ][i % TQ];
}

// ------------------------------------------------------------------ path product + top-k
constexpr int PT_THREADS = 256, PT_CHUNK = 4096;

struct Cand {
    float s;
    int sid;
};
// order: score desc, then sentence id asc; sid < 0 = empty
__device__ __forceinline__ bool cand_better(float as, int ai, float bs, int bi) {
    if (bi < 0) return ai >= 0;
    if (ai < 0) return false;
    if (as != bs) return as > bs;
    return ai < bi;
}

// Repeated block arg-max over `n` candidates held in shared memory; writes the k best in order.
__device__ void block_select_topk(float *cs, int *ci, int n, int k, float *out_s, int *out_i, float *ws, int *wi,
                                  int *wp) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    for (int r = 0; r < k; r++) {
        float bs = 0.f;
        int bi = -1, bp = -1;
        for (int i = tid; i < n; i += blockDim.x) {
            if (cand_better(cs[i], ci[i], bs, bi)) { bs = cs[i]; bi = ci[i]; bp = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            float os = __shfl_xor_sync(0xffffffffu, bs, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o), op = __shfl_xor_sync(0xffffffffu, bp, o);
            if (cand_better(os, oi, bs, bi)) { bs = os; bi = oi; bp = op; }
        }
        if (lane == 0) { ws[warp] = bs; wi[warp] = bi; wp[warp] = bp; }
        __syncthreads();
        if (warp == 0) {
            bs = 0.f; bi = -1; bp = -1;
            if (lane < nw) { bs = ws[lane]; bi = wi[lane]; bp = wp[lane]; }
            for (int o = 16; o > 0; o >>= 1) {
                float os = __shfl_xor_sync(0xffffffffu, bs, o);
                int oi = __shfl_xor_sync(0xffffffffu, bi, o), op = __shfl_xor_sync(0xffffffffu, bp, o);
                if (cand_better(os, oi, bs, bi)) { bs = os; bi = oi; bp = op; }
            }
            if (lane == 0) {
                out_s[r] = bs;
                out_i[r] = bi;
                if (bp >= 0) ci[bp] = -1;  // remove the winner
            }
        }
        __syncthreads();
    }
}

// grid (chunks, nq): leaf scores of one chunk of positions for one query + its k best
__global__ void __launch_bounds__(PT_THREADS)
paths_topk_kernel(const float *__restrict__ node_scores, long long ld, int n_pos, int max_len,
                  const int *__restrict__ path_idx, const float *__restrict__ path_w,
                  const int *__restrict__ pos_sid, int k, float *leaf_scores, float *cand_s, int *cand_i,
                  int n_chunks) {
    __shared__ float cs[PT_CHUNK];
    __shared__ int ci[PT_CHUNK];
    __shared__ float ws[32], os_[CW_MAX_K];
    __shared__ int wi[32], wp[32], oi_[CW_MAX_K];
    const int chunk = blockIdx.x;
    const long long q = blockIdx.y;
    const float *s = node_scores + q * ld;
    const int p0 = chunk * PT_CHUNK;
    const int n = min(PT_CHUNK, n_pos - p0);
    for (int i = threadIdx.x; i < n; i += PT_THREADS) {
        const int p = p0 + i;
        float acc = 0.0f;
        for (int j = 0; j < max_len; j++) {
            const int b = path_idx[(size_t)j * n_pos + p];
            if (b < 0) break;
            acc = __fmaf_rn(path_w[(size_t)j * n_pos + p], s[b], acc);
        }
        const int sid = pos_sid[p];
        cs[i] = acc;
        ci[i] = sid;
        if (leaf_scores) leaf_scores[q * n_pos + sid] = acc;
    }
    __syncthreads();
    if (k > 0) {
        block_select_topk(cs, ci, n, k, os_, oi_, ws, wi, wp);
        for (int r = threadIdx.x; r < k; r += PT_THREADS) {
            cand_s[(q * n_chunks + chunk) * k + r] = os_[r];
            cand_i[(q * n_chunks + chunk) * k + r] = oi_[r];
        }
    }
}

// grid (nq): merge n_chunks*k candidates of one query into the final k
__global__ void __launch_bounds__(PT_THREADS)
merge_topk_kernel(const float *__restrict__ cand_s, const int *__restrict__ cand_i, int n_chunks, int k,
                  int *out_sid, float *out_score) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int n = n_chunks * k;
    float *cs = reinterpret_cast<float *>(sm_raw);
    int *ci = reinterpret_cast<int *>(cs + n);
    __shared__ float ws[32], os_[CW_MAX_K];
    __shared__ int wi[32], wp[32], oi_[CW_MAX_K];
    const long long q = blockIdx.x;
    for (int i = threadIdx.x; i < n; i += PT_THREADS) {
        cs[i] = cand_s[q * n + i];
        ci[i] = cand_i[q * n + i];
    }
    __syncthreads();
    block_select_topk(cs, ci, n, k, os_, oi_, ws, wi, wp);
    for (int r = threadIdx.x; r < k; r += PT_THREADS) {
        out_sid[q * k + r] = oi_[r];
        out_score[q * k + r] = oi_[r] >= 0 ? os_[r] : -__int_as_float(0x7f800000);
    }
}

}  // namespace cw

using namespace cw;

static int score_smem_bytes() { return STAGES * (int)sizeof(ScoreStage) + 2 * STAGES * (int)sizeof(uint64_t); }

extern "C" int64_t cw_xt_floats(int64_t nq, int32_t D) {
    return ((nq + TQ - 1) / TQ) * (int64_t)((D + TK - 1) / TK) * TK * TQ;
}

extern "C" int cw_dense_node_scores(const cw_index *ix, const float *Q, int64_t nq, float *xt_scratch,
                                    float *node_scores, int64_t ld, void *stream) {
    if (!ix || !Q || !node_scores || !xt_scratch || nq < 0 || ld < (int64_t)ix->n_ntiles * TN || (ld & 3)) {
        cw_set_error("cw_dense_node_scores: bad argument (ld must be >= n_ntiles*%d and a multiple of 4)", TN);
        return CW_E_ARG;
    }
    if (nq == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int n_qtiles = (int)((nq + TQ - 1) / TQ);
    int rc = 0;
    static bool configured = false;
    if (!configured) {
        rc = cw_check_cuda(cudaFuncSetAttribute(dense_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                score_smem_bytes()),
                           "cw_dense: smem attribute");
        if (rc) return rc;
        configured = true;
    }
    tile_queries_kernel<<<dim3(n_qtiles, ix->n_ktiles), 256, 0, st>>>(Q, nq, ix->D, ix->n_ktiles, xt_scratch);
    dense_score_kernel<<<dim3(n_qtiles, ix->n_ntiles), SCORE_THREADS, score_smem_bytes(), st>>>(
        xt_scratch, ix->R, ix->MB, ix->sumlog, node_scores, ld, nq, ix->n_ktiles);
    return cw_check_cuda(cudaGetLastError(), "cw_dense_node_scores");
}

extern "C" int64_t cw_topk_chunks(int64_t n_pos) { return (n_pos + PT_CHUNK - 1) / PT_CHUNK; }

extern "C" int cw_dense_paths_topk(const cw_index *ix, const float *node_scores, int64_t ld, int64_t nq, int k,
                                   float *leaf_scores, int32_t *out_sid, float *out_score, int32_t *scratch,
                                   void *stream) {
    if (!ix || !node_scores || nq < 0 || k < 0 || k > CW_MAX_K || ix->n_pos < 1 || !ix->path_idx || !ix->path_w ||
        !ix->pos_sid || (k > 0 && (!out_sid || !out_score || !scratch))) {
        cw_set_error("cw_dense_paths_topk: bad argument (k=%d, max %d)", k, CW_MAX_K);
        return CW_E_ARG;
    }
    if (nq == 0) return 0;
    if (nq > 65535) {
        cw_set_error("cw_dense_paths_topk: at most 65535 queries per call (got %lld)", (long long)nq);
        return CW_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int n_chunks = (int)cw_topk_chunks(ix->n_pos);
    float *cand_s = reinterpret_cast<float *>(scratch);
    int *cand_i = scratch + (size_t)nq * n_chunks * (k > 0 ? k : 1);
    paths_topk_kernel<<<dim3(n_chunks, (unsigned)nq), PT_THREADS, 0, st>>>(
        node_scores, ld, ix->n_pos, ix->max_len, ix->path_idx, ix->path_w, ix->pos_sid, k, leaf_scores, cand_s,
        cand_i, n_chunks);
    if (k > 0) {
        size_t smem = (size_t)n_chunks * k * 8;
        if (smem > 48 * 1024) {
            int rc = cw_check_cuda(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                        (int)smem),
                                   "cw_dense_paths_topk: smem attribute");
            if (rc) return rc;
        }
        merge_topk_kernel<<<(unsigned)nq, PT_THREADS, smem, st>>>(cand_s, cand_i, n_chunks, k, out_sid, out_score);
    }
    return cw_check_cuda(cudaGetLastError(), "cw_dense_paths_topk");
}

extern "C" int cw_predict_dense_host(const cw_index *ix, const float *Q_host, int64_t nq, int k, float *Q_dev,
                                     float *xt_scratch, float *node_scores, int64_t ld, int32_t *out_sid_dev, float *out_score_dev,
                                     int32_t *scratch, int32_t *out_sid_host, float *out_score_host, void *stream) {
    if (!ix || !Q_host || !Q_dev || !out_sid_host || !out_score_host || k < 1) {
        cw_set_error("cw_predict_dense_host: bad argument");
        return CW_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = cw_check_cuda(cudaMemcpyAsync(Q_dev, Q_host, (size_t)nq * ix->D * sizeof(float), cudaMemcpyHostToDevice, st),
                           "cw_predict_dense_host: H2D");
    if (rc) return rc;
    if ((rc = cw_dense_node_scores(ix, Q_dev, nq, xt_scratch, node_scores, ld, stream))) return rc;
    if ((rc = cw_dense_paths_topk(ix, node_scores, ld, nq, k, nullptr, out_sid_dev, out_score_dev, scratch, stream)))
        return rc;
    rc = cw_check_cuda(cudaMemcpyAsync(out_sid_host, out_sid_dev, (size_t)nq * k * sizeof(int32_t),
                                       cudaMemcpyDeviceToHost, st),
                       "cw_predict_dense_host: D2H ids");
    if (rc) return rc;
    rc = cw_check_cuda(cudaMemcpyAsync(out_score_host, out_score_dev, (size_t)nq * k * sizeof(float),
                                       cudaMemcpyDeviceToHost, st),
                       "cw_predict_dense_host: D2H scores");
    if (rc) return rc;
    return cw_check_cuda(cudaStreamSynchronize(st), "cw_predict_dense_host: sync");
}
