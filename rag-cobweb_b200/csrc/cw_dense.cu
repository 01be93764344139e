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

// packed fp32x2 arithmetic (Blackwell FFMA2): one instruction = two FMAs, operands are aligned
// 64-bit register pairs, so three-operand FMAs no longer collide in the two register banks.
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// Refill protocol: every warp bumps a per-stage counter when it is done with the stage; the
// warp that arrives last re-arms the stage's mbarrier and issues the three bulk copies for the
// k-tile STAGES ahead.  No thread ever blocks waiting for the others (v1 had thread 0 wait on
// an "empty" mbarrier, which put 35 % of warp time into mbarrier waits -- profiles/r01_*).
//
// grid: (query tiles, node tiles) -- query tile fastest so that concurrently resident CTAs
// share one node tile through L2 and the node matrices stream from HBM exactly once.
// Warp w owns a 32-query x 64-node sub-tile (4 x 2 warps); lane (tx = lane%8, ty = lane/8)
// owns queries {ty*4.., 16+ty*4..} x nodes {tx*4.., 32+tx*4..}: every LDS.128 touches one
// 128-byte line per half-warp (2 wavefronts instead of 4).
__global__ void __launch_bounds__(SCORE_THREADS, 2)
dense_score_kernel(const float *__restrict__ XT, const float *__restrict__ R, const float *__restrict__ MB,
                   const float *__restrict__ sumlog, float *__restrict__ out, long long ld, long long nq,
                   int n_ktiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ScoreStage *st = reinterpret_cast<ScoreStage *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + STAGES * sizeof(ScoreStage));
    int *done = reinterpret_cast<int *>(full + STAGES);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qt = blockIdx.x, nt = blockIdx.y;
    const float *xsrc = XT + (size_t)qt * n_ktiles * (TK * TQ);
    const float *rsrc = R + (size_t)nt * n_ktiles * (TK * TN);
    const float *msrc = MB + (size_t)nt * n_ktiles * (TK * TN);
    constexpr uint32_t XB = TK * TQ * 4, NB = TK * TN * 4;
    constexpr int NWARPS = SCORE_THREADS / 32;

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);
            done[s] = 0;
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

    const int tx = lane & 7, ty = lane >> 3;
    const int qb = (warp >> 1) * 32 + ty * 4;  // first query of this thread inside the tile (second group: +16)
    const int nb = (warp & 1) * 64 + tx * 4;   // first node (second group: +32)
    uint64_t acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int p = 0; p < 4; p++) acc[i][p] = 0ull;

    for (int kt = 0; kt < n_ktiles; kt++) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        mbar_wait(&full[s], ph);
        const ScoreStage &S = st[s];
#pragma unroll 4
        for (int kk = 0; kk < TK; kk++) {
            const float4 xa = *reinterpret_cast<const float4 *>(&S.x[kk][qb]);
            const float4 xb = *reinterpret_cast<const float4 *>(&S.x[kk][qb + 16]);
            const ulonglong2 ra = *reinterpret_cast<const ulonglong2 *>(&S.r[kk][nb]);
            const ulonglong2 rb = *reinterpret_cast<const ulonglong2 *>(&S.r[kk][nb + 32]);
            const ulonglong2 ma = *reinterpret_cast<const ulonglong2 *>(&S.mb[kk][nb]);
            const ulonglong2 mb = *reinterpret_cast<const ulonglong2 *>(&S.mb[kk][nb + 32]);
            const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            const uint64_t r2[4] = {ra.x, ra.y, rb.x, rb.y};
            const uint64_t m2[4] = {ma.x, ma.y, mb.x, mb.y};
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint64_t xx = pack2(xv[i], xv[i]);
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const uint64_t u = ffma2(xx, r2[p], m2[p]);
                    acc[i][p] = ffma2(u, u, acc[i][p]);
                }
            }
        }
        // release the stage; the last warp to get here refills it
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            const int prev = atomicAdd(&done[s], 1);
            if (prev == NWARPS - 1) {
                atomicExch(&done[s], 0);
                const int k2 = kt + STAGES;
                if (k2 < n_ktiles) {
                    __threadfence_block();
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive_expect_tx(&full[s], XB + 2 * NB);
                    bulk_g2s(st[s].x, xsrc + (size_t)k2 * (TK * TQ), XB, &full[s]);
                    bulk_g2s(st[s].r, rsrc + (size_t)k2 * (TK * TN), NB, &full[s]);
                    bulk_g2s(st[s].mb, msrc + (size_t)k2 * (TK * TN), NB, &full[s]);
                }
            }
        }
    }

    // epilogue: -0.5 * (sumlog + quad)   (CobwebWrapper.py:232-236)
    const int b0 = nt * TN + nb;
    const float4 sla = *reinterpret_cast<const float4 *>(sumlog + b0);
    const float4 slb = *reinterpret_cast<const float4 *>(sumlog + b0 + 32);
    const float sl[8] = {sla.x, sla.y, sla.z, sla.w, slb.x, slb.y, slb.z, slb.w};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const long long q = (long long)qt * TQ + qb + (i < 4 ? i : 16 + (i - 4));
        if (q < nq) {
            float o[8];
#pragma unroll
            for (int p = 0; p < 4; p++) {
                float lo, hi;
                unpack2(acc[i][p], lo, hi);
                o[2 * p] = -0.5f * (sl[2 * p] + lo);
                o[2 * p + 1] = -0.5f * (sl[2 * p + 1] + hi);
            }
            float *row = out + q * ld + b0;
            *reinterpret_cast<float4 *>(row) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4 *>(row + 32) = make_float4(o[4], o[5], o[6], o[7]);
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
constexpr int PT_THREADS = 256, PT_CHUNK = 1024, PT_QB = 8, PT_MAXLEN = 64;

// order: score desc, then sentence id asc; sid < 0 = empty
__device__ __forceinline__ bool cand_better(float as, int ai, float bs, int bi) {
    if (bi < 0) return ai >= 0;
    if (ai < 0) return false;
    if (as != bs) return as > bs;
    return ai < bi;
}

// One warp selects the k best of n (score, sid) candidates held in shared memory, in order.
// k <= 64: single pass.  The running top-k is kept sorted across the warp (rank r lives in
// lane r%32, slot r/32); each step tests 32 candidates against the current k-th best with one
// ballot and inserts the few that beat it (expected k*ln(n/k) insertions in total).
// k > 64: k rounds of arg-max + knock-out.
__device__ __forceinline__ void warp_select_topk(float *cs, const int *ci, int n, int k, float *out_s, int *out_i) {
    const int lane = threadIdx.x & 31;
    const float NEG_INF = -__int_as_float(0x7f800000);
    if (k <= 64) {
        float ls[2] = {NEG_INF, NEG_INF};  // sorted list, best first
        int li[2] = {-1, -1};
        const int last_lane = (k - 1) & 31, last_slot = (k - 1) >> 5;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            float v = NEG_INF;
            int id = -1;
            if (i < n) { v = cs[i]; id = ci[i]; }
            // current k-th best (threshold)
            const float ts = __shfl_sync(0xffffffffu, last_slot ? ls[1] : ls[0], last_lane);
            const int ti = __shfl_sync(0xffffffffu, last_slot ? li[1] : li[0], last_lane);
            unsigned m = __ballot_sync(0xffffffffu, cand_better(v, id, ts, ti));
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const float cv = __shfl_sync(0xffffffffu, v, src);
                const int cid = __shfl_sync(0xffffffffu, id, src);
                // rank of the candidate = number of list entries that beat it
                const unsigned b0 = __ballot_sync(0xffffffffu, cand_better(ls[0], li[0], cv, cid));
                const unsigned b1 = __ballot_sync(0xffffffffu, cand_better(ls[1], li[1], cv, cid));
                const int pos = __popc(b0) + __popc(b1);
                if (pos >= k) continue;  // an earlier insertion of this step raised the bar
                // shift entries at rank >= pos down by one, slot 1 first (it takes lane 31 of slot 0)
                const float up0s = __shfl_up_sync(0xffffffffu, ls[0], 1), up1s = __shfl_up_sync(0xffffffffu, ls[1], 1);
                const int up0i = __shfl_up_sync(0xffffffffu, li[0], 1), up1i = __shfl_up_sync(0xffffffffu, li[1], 1);
                const float carry_s = __shfl_sync(0xffffffffu, ls[0], 31);
                const int carry_i = __shfl_sync(0xffffffffu, li[0], 31);
                const int r0 = lane, r1 = 32 + lane;
                if (r1 > pos) { ls[1] = lane == 0 ? carry_s : up1s; li[1] = lane == 0 ? carry_i : up1i; }
                if (r1 == pos) { ls[1] = cv; li[1] = cid; }
                if (r0 > pos) { ls[0] = up0s; li[0] = up0i; }
                if (r0 == pos) { ls[0] = cv; li[0] = cid; }
            }
        }
        if (lane < k) { out_s[lane] = ls[0]; out_i[lane] = li[0]; }
        if (32 + lane < k) { out_s[32 + lane] = ls[1]; out_i[32 + lane] = li[1]; }
        __syncwarp();
        return;
    }
    for (int r = 0; r < k; r++) {
        float bs = 0.f;
        int bi = -1, bp = -1;
        for (int i = lane; i < n; i += 32) {
            const float v = cs[i];
            const int id = (v == NEG_INF) ? -1 : ci[i];
            if (cand_better(v, id, bs, bi)) { bs = v; bi = id; bp = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o), op = __shfl_xor_sync(0xffffffffu, bp, o);
            if (cand_better(os, oi, bs, bi)) { bs = os; bi = oi; bp = op; }
        }
        if (lane == 0) {
            out_s[r] = bs;
            out_i[r] = bi;
            if (bp >= 0) cs[bp] = NEG_INF;  // knock out the winner (scores are finite)
        }
        __syncwarp();
    }
}

// grid (chunks, query blocks of PT_QB): leaf scores of one chunk of positions for PT_QB queries
// (path indices are read once per block of queries), then warp q picks query q's k best.
// leaf = sequential FMA over the path, root first, of w_table[len][j] * node score.
__global__ void __launch_bounds__(PT_THREADS)
paths_topk_kernel(const float *__restrict__ node_scores, long long ld, long long nq, int n_pos, int max_len,
                  const int *__restrict__ path_idx, const int *__restrict__ path_len,
                  const float *__restrict__ w_table, const int *__restrict__ pos_sid, int k, float *leaf_scores,
                  float *cand_s, int *cand_i, int n_chunks) {
    extern __shared__ __align__(16) unsigned char pt_smem[];
    float *cs = reinterpret_cast<float *>(pt_smem);             // [PT_QB][PT_CHUNK]
    int *ci = reinterpret_cast<int *>(cs + PT_QB * PT_CHUNK);   // [PT_CHUNK] sentence ids
    float *wt = reinterpret_cast<float *>(ci + PT_CHUNK);       // [(max_len+1) * max_len]
    float *os_ = wt + (PT_MAXLEN + 1) * PT_MAXLEN;              // [PT_QB][CW_MAX_K]
    int *oi_ = reinterpret_cast<int *>(os_ + PT_QB * CW_MAX_K);

    const int chunk = blockIdx.x;
    const long long q0 = (long long)blockIdx.y * PT_QB;
    const int nqb = (int)min((long long)PT_QB, nq - q0);
    const int p0 = chunk * PT_CHUNK;
    const int n = min(PT_CHUNK, n_pos - p0);
    for (int i = threadIdx.x; i < (max_len + 1) * max_len; i += PT_THREADS) wt[i] = w_table[i];
    __syncthreads();
    const float *s0 = node_scores + q0 * ld;
    for (int i = threadIdx.x; i < n; i += PT_THREADS) {
        const int p = p0 + i;
        const int len = path_len[p];
        const float *w = wt + len * max_len;
        float acc[PT_QB];
#pragma unroll
        for (int q = 0; q < PT_QB; q++) acc[q] = 0.0f;
        // 4 path levels at a time: all index loads, then all 32 gathers, then the FMA chains in
        // path order (keeps ~32 independent loads in flight instead of one dependent pair)
        for (int j0 = 0; j0 < len; j0 += 4) {
            int b[4];
#pragma unroll
            for (int u = 0; u < 4; u++) b[u] = (j0 + u < len) ? path_idx[(size_t)(j0 + u) * n_pos + p] : -1;
            float sv[4][PT_QB];
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int q = 0; q < PT_QB; q++) sv[u][q] = (b[u] >= 0 && q < nqb) ? s0[(size_t)q * ld + b[u]] : 0.0f;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (b[u] >= 0) {
                    const float wj = w[j0 + u];
#pragma unroll
                    for (int q = 0; q < PT_QB; q++) acc[q] = __fmaf_rn(wj, sv[u][q], acc[q]);
                }
            }
        }
        const int sid = pos_sid[p];
        ci[i] = sid;
#pragma unroll
        for (int q = 0; q < PT_QB; q++) {
            cs[q * PT_CHUNK + i] = acc[q];
            if (leaf_scores && q < nqb) leaf_scores[(q0 + q) * n_pos + sid] = acc[q];
        }
    }
    __syncthreads();
    if (k > 0) {
        const int q = threadIdx.x >> 5;
        if (q < nqb) {
            warp_select_topk(cs + q * PT_CHUNK, ci, n, k, os_ + q * CW_MAX_K, oi_ + q * CW_MAX_K);
            for (int r = threadIdx.x & 31; r < k; r += 32) {
                cand_s[((q0 + q) * n_chunks + chunk) * k + r] = os_[q * CW_MAX_K + r];
                cand_i[((q0 + q) * n_chunks + chunk) * k + r] = oi_[q * CW_MAX_K + r];
            }
        }
    }
}

// grid (ceil(nq/8)): warp w merges the n_chunks*k candidates of one query into the final k
__global__ void __launch_bounds__(PT_THREADS)
merge_topk_kernel(const float *__restrict__ cand_s, const int *__restrict__ cand_i, long long nq, int n_chunks, int k,
                  int *out_sid, float *out_score) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int n = n_chunks * k;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *cs = reinterpret_cast<float *>(sm_raw) + (size_t)w * n;
    int *ci = reinterpret_cast<int *>(reinterpret_cast<float *>(sm_raw) + (size_t)(PT_THREADS / 32) * n) + (size_t)w * n;
    __shared__ float os_[PT_THREADS / 32][CW_MAX_K];
    __shared__ int oi_[PT_THREADS / 32][CW_MAX_K];
    const long long q = (long long)blockIdx.x * (PT_THREADS / 32) + w;
    if (q >= nq) return;
    const float NEG_INF = -__int_as_float(0x7f800000);
    for (int i = lane; i < n; i += 32) {
        const int id = cand_i[q * n + i];
        cs[i] = id >= 0 ? cand_s[q * n + i] : NEG_INF;
        ci[i] = id;
    }
    __syncwarp();
    warp_select_topk(cs, ci, n, k, os_[w], oi_[w]);
    for (int r = lane; r < k; r += 32) {
        out_sid[q * k + r] = oi_[w][r];
        out_score[q * k + r] = oi_[w][r] >= 0 ? os_[w][r] : NEG_INF;
    }
}

}  // namespace cw

using namespace cw;

static int score_smem_bytes() { return STAGES * (int)sizeof(ScoreStage) + STAGES * (int)(sizeof(uint64_t) + sizeof(int)); }

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

static size_t paths_smem_bytes() {
    return (size_t)PT_QB * PT_CHUNK * 4 + (size_t)PT_CHUNK * 4 + (size_t)(PT_MAXLEN + 1) * PT_MAXLEN * 4 +
           (size_t)PT_QB * CW_MAX_K * 8;
}

extern "C" int cw_dense_paths_topk(const cw_index *ix, const float *node_scores, int64_t ld, int64_t nq, int k,
                                   float *leaf_scores, int32_t *out_sid, float *out_score, int32_t *scratch,
                                   void *stream) {
    if (!ix || !node_scores || nq < 0 || k < 0 || k > CW_MAX_K || ix->n_pos < 1 || !ix->path_idx || !ix->path_len ||
        !ix->w_table || !ix->pos_sid || ix->max_len < 1 || ix->max_len > PT_MAXLEN ||
        (k > 0 && (!out_sid || !out_score || !scratch))) {
        cw_set_error("cw_dense_paths_topk: bad argument (k=%d max %d, max_len=%d max %d)", k, CW_MAX_K,
                     ix ? ix->max_len : -1, PT_MAXLEN);
        return CW_E_ARG;
    }
    if (nq == 0) return 0;
    if (nq > 65535LL * PT_QB) {
        cw_set_error("cw_dense_paths_topk: at most %d queries per call (got %lld)", 65535 * PT_QB, (long long)nq);
        return CW_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int n_chunks = (int)cw_topk_chunks(ix->n_pos);
    float *cand_s = reinterpret_cast<float *>(scratch);
    int *cand_i = scratch + (size_t)nq * n_chunks * (k > 0 ? k : 1);
    static bool configured = false;
    if (!configured) {
        int rc = cw_check_cuda(cudaFuncSetAttribute(paths_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    (int)paths_smem_bytes()),
                               "cw_dense_paths_topk: smem attribute");
        if (rc) return rc;
        configured = true;
    }
    const unsigned qblocks = (unsigned)((nq + PT_QB - 1) / PT_QB);
    paths_topk_kernel<<<dim3(n_chunks, qblocks), PT_THREADS, paths_smem_bytes(), st>>>(
        node_scores, ld, nq, ix->n_pos, ix->max_len, ix->path_idx, ix->path_len, ix->w_table, ix->pos_sid, k,
        leaf_scores, cand_s, cand_i, n_chunks);
    if (k > 0) {
        const int wpb = PT_THREADS / 32;
        size_t smem = (size_t)wpb * n_chunks * k * 8;
        if (smem > 200 * 1024) {
            cw_set_error("cw_dense_paths_topk: k=%d with %d position chunks needs %zu bytes of shared memory", k, n_chunks, smem);
            return CW_E_ARG;
        }
        static size_t merge_configured = 48 * 1024;
        if (smem > merge_configured) {
            int rc = cw_check_cuda(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                        (int)smem),
                                   "cw_dense_paths_topk: smem attribute");
            if (rc) return rc;
            merge_configured = smem;
        }
        merge_topk_kernel<<<(unsigned)((nq + wpb - 1) / wpb), PT_THREADS, smem, st>>>(cand_s, cand_i, nq, n_chunks, k,
                                                                                    out_sid, out_score);
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
