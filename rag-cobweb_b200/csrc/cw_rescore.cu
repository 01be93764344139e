// cw_rescore.cu -- exact re-score of the tensor-core pre-filter (cw_tensor.cu).
//
// "tf32x3" dense predict = tcgen05 node scores -> path product -> top-kc candidates per query
// (kc > k), then this kernel recomputes the leaf scores of the candidates with EXACTLY the
// arithmetic of the FP32-pipe path (dense_score_kernel + paths_topk_kernel in cw_dense.cu):
//   node:  acc = 0; for d ascending: u = fma(x_d, r_d, mb_d); acc = fma(u, u, acc);
//          s = -0.5f * (sumlog + acc)
//   leaf:  acc = 0; root first: acc = fma((float)(level_w[j]/len), s_j, acc)
// and keeps the best k by (score desc, sentence id asc).  With eps a bound of |approximate - exact|
// for any leaf score of the query and a_k the k-th best approximate score, a leaf whose approximate
// score is below a_k - 2 eps cannot be in the exact top-k (the k approximate leaders all beat it
// exactly), so only the m candidates at or above that threshold are re-scored; if all kc candidates
// are (m == kc) the list may be incomplete and the query is flagged for the FP32 path instead.
// Result: ids and scores bit-identical to the "fp32" mode (tests/test_gpu_parity.py asserts equality).
// eps is a statistical bound -- operand term 2^-18 T plus (4 + 3 sqrt(max_len)) ulps of the score magnitude for the
// roundings of the two FMA chains -- 4.7x the largest deviation seen over 1.3e9 leaf scores at cfg3 (depth 14) and
// 5e8 at cfg4 (depth 66), not a worst-case proof (a worst-case fp32 accumulation bound is useless for either kernel).
//
// Strict arithmetic (-fmad=false; FMAs are explicit): r / mb come from a row-major copy of the
// index operands built here with the same operations cw_index.cu uses for the index tiles.
#include "cw_common.cuh"

void cw_set_error(const char *fmt, ...);
int cw_check_cuda(cudaError_t e, const char *what);

namespace cw {

constexpr int RS_THREADS = 128, RS_WARPS = RS_THREADS / 32;

struct RescoreArgs {
    cw_store s;
    cw_index ix;
    const float2 *RM;       // [nn, D] {r, mb} per index row, row-major
    const int *pos_of_sid;  // sentence id -> position
    const float *Q;
    long long nq;
    int kc, k;
    const int *cand_sid;       // [nq, kc] approximate top-kc, best first, -1 padded
    const float *cand_score;   // [nq, kc]
    float inv_prior, hmax, lmax, wfac, eps_scale, mag_scale;
    int *out_sid;       // [nq, k]
    float *out_score;   // [nq, k]
    int *fail;          // [1 + nq]: count, then the flagged queries
};

__global__ void __launch_bounds__(RS_THREADS)
rescore_kernel(const RescoreArgs a) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    const int D = a.s.D, ML = a.ix.max_len, kc = a.kc, EMAX = kc * ML;
    double *lw = reinterpret_cast<double *>(rs_smem);                    // [ML]
    float2 *stage = reinterpret_cast<float2 *>(lw + ML);                 // [RS_WARPS][32][33]
    float *xq = reinterpret_cast<float *>(stage + RS_WARPS * 32 * 33);   // [D]
    int *nid = reinterpret_cast<int *>(xq + D);                          // [E] index row of (candidate, level)
    int *ulist = nid + EMAX;                                                // [E] unique rows
    float *uscore = reinterpret_cast<float *>(ulist + EMAX);                // [E]
    int *cp = reinterpret_cast<int *>(uscore + EMAX);                       // [kc] position
    int *clen = cp + kc;                                                 // [kc]
    int *csid = clen + kc;                                               // [kc]
    float *cex = reinterpret_cast<float *>(csid + kc);                   // [kc] exact leaf score
    float *red = cex + kc;                                               // [RS_WARPS]
    int *ucount = reinterpret_cast<int *>(red + RS_WARPS);               // [1]
    unsigned short *firstc = reinterpret_cast<unsigned short *>(ucount + 1);  // [E]
    unsigned short *slot = firstc + EMAX;                                   // [E]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int4 *pos_rec = reinterpret_cast<const int4 *>(a.ix.pos_rec);
    for (int i = tid; i < ML; i += RS_THREADS) lw[i] = a.ix.level_w[i];

    for (long long q = blockIdx.x; q < a.nq; q += gridDim.x) {
        __syncthreads();
        // ---- query row, |x|^2
        float xx = 0.0f;
        for (int d = tid; d < D; d += RS_THREADS) {
            const float v = a.Q[q * D + d];
            xq[d] = v;
            xx = fmaf(v, v, xx);
        }
        for (int off = 16; off > 0; off >>= 1) xx += __shfl_xor_sync(0xffffffffu, xx, off);
        if (lane == 0) red[warp] = xx;
        if (tid == 0) *ucount = 0;
        __syncthreads();
        // ---- which candidates can still be in the exact top-k
        float x2 = 0.0f;
        for (int w = 0; w < RS_WARPS; w++) x2 += red[w];
        const float tq = 2.0f * (x2 * a.inv_prior + a.hmax);
        const float eps = a.wfac * (a.eps_scale * tq + a.mag_scale * (4.0f + 3.0f * sqrtf((float)ML)) * 0.5f * (a.lmax + a.hmax + tq));
        const float *cs = a.cand_score + q * kc;
        const int *ci = a.cand_sid + q * kc;
        int nvalid = 0, m = 0;
        {
            const int kth = ci[a.k - 1] >= 0 ? a.k - 1 : -1;  // fewer than k sentences: keep every valid candidate
            const float thr = kth >= 0 ? cs[kth] - 2.0f * eps : -__int_as_float(0x7f800000);
            for (int c = 0; c < kc; c++) {
                const bool valid = ci[c] >= 0;
                nvalid += valid;
                m += valid && cs[c] >= thr;  // the list is sorted best first, so these are the first m
            }
        }
        if (tid == 0 && m == kc) {  // even the weakest candidate is within 2 eps: the list may be incomplete
            const int at = atomicAdd(a.fail, 1);
            a.fail[1 + at] = (int)q;
        }
        if (tid < m) {
            const int sid = ci[tid];
            const int p = a.pos_of_sid[sid];
            csid[tid] = sid;
            cp[tid] = p;
            clen[tid] = pos_rec[p].x;
        }
        __syncthreads();
        // ---- (candidate, level) -> index row; unique rows get a slot
        const int E = m * ML;
        for (int e = tid; e < E; e += RS_THREADS) {
            const int c = e / ML, j = e - c * ML;
            nid[e] = j < clen[c] ? a.ix.path_idx[(size_t)cp[c] * ML + j] : -1;
        }
        __syncthreads();
        for (int e = tid; e < E; e += RS_THREADS) {
            const int b = nid[e];
            if (b < 0) continue;
            const int c = e / ML, j = e - c * ML;
            int f = 0;
            while (nid[f * ML + j] != b) f++;  // first candidate through this node (f <= c)
            firstc[e] = (unsigned short)f;
            if (f == c) {
                const int sl = atomicAdd(ucount, 1);
                ulist[sl] = b;
                slot[e] = (unsigned short)sl;
            }
        }
        __syncthreads();
        for (int e = tid; e < E; e += RS_THREADS) {
            if (nid[e] < 0) continue;
            const int c = e / ML, j = e - c * ML, f = firstc[e];
            if (f != c) slot[e] = slot[f * ML + j];
        }
        const int U = *ucount;
        // ---- exact node scores: the unique rows are dealt evenly to the warps, lane = row.  Per 32 attributes the
        // warp reads its rows with coalesced 256-byte loads (all in flight at once, and one segment ahead of the
        // arithmetic), transposes them through shared memory, and every lane runs its row's FMA chain in order.
        float2 *stw = stage + warp * 32 * 33;
        for (int base0 = 0; base0 < U; base0 += 32 * RS_WARPS) {
            const int in_round = min(U - base0, 32 * RS_WARPS);
            const int per = (in_round + RS_WARPS - 1) / RS_WARPS;
            const int my_lo = base0 + warp * per;
            const int my_n = max(0, min(per, base0 + in_round - my_lo));
            if (my_n == 0) continue;  // warp-uniform
            const int u = my_lo + lane;
            const int b = lane < my_n ? ulist[u] : -1;
            float acc = 0.0f;
            float2 o[32];
            auto fetch = [&](int d0) {
                const int d = d0 + lane;
#pragma unroll
                for (int r = 0; r < 32; r++) {
                    const int rb = __shfl_sync(0xffffffffu, b, r);
                    o[r] = make_float2(0.0f, 0.0f);
                    if (rb >= 0 && d < D) o[r] = a.RM[(size_t)rb * D + d];
                }
            };
            fetch(0);
            for (int d0 = 0; d0 < D; d0 += 32) {
#pragma unroll
                for (int r = 0; r < 32; r++) stw[r * 33 + lane] = o[r];
                __syncwarp();
                if (d0 + 32 < D) fetch(d0 + 32);
                const int nd = min(32, D - d0);
                for (int j = 0; j < nd; j++) {
                    const float2 v = stw[lane * 33 + j];
                    const float t = __fmaf_rn(xq[d0 + j], v.x, v.y);
                    acc = __fmaf_rn(t, t, acc);
                }
                __syncwarp();
            }
            if (b >= 0) uscore[u] = -0.5f * (a.ix.sumlog[b] + acc);
        }
        __syncthreads();
        // ---- exact leaf scores of the candidates
        if (tid < m) {
            const int len = clen[tid];
            float acc = 0.0f;
            for (int j = 0; j < len; j++) {
                const float wl = (float)(lw[j] / (double)len);
                acc = __fmaf_rn(wl, uscore[slot[tid * ML + j]], acc);
            }
            cex[tid] = acc;
        }
        __syncthreads();
        // ---- rank by (score desc, sentence id asc)
        if (tid < m) {
            const int sid = csid[tid];
            const float v = cex[tid];
            int rank = 0;
            for (int c = 0; c < m; c++) {
                const float ov = cex[c];
                if (ov > v || (ov == v && csid[c] < sid)) rank++;
            }
            if (rank < a.k) {
                a.out_sid[q * a.k + rank] = sid;
                a.out_score[q * a.k + rank] = v;
            }
        }
        for (int r = min(nvalid, m) + tid; r < a.k; r += RS_THREADS) {  // fewer than k sentences: empty ranks
            a.out_sid[q * a.k + r] = -1;
            a.out_score[q * a.k + r] = -__int_as_float(0x7f800000);
        }
    }
}

// RM[b, d] = {r, mb} of index row b (node order[b]), row-major, the operands of the FP32-pipe score
__global__ void __launch_bounds__(256)
rescore_rows_kernel(cw_store s, const int *__restrict__ order, int nn, float2 *RM) {
    const int D = s.D;
    const bool cutoff = (s.flags & CW_ACUITY_CUTOFF) != 0;
    const float prior = s.prior_var;
    for (int b = blockIdx.x; b < nn; b += gridDim.x) {
        const int node = order[b];
        const float cnt = s.count[node];
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            float r, mb;
            dense_operands(s.mean[(size_t)node * D + d], s.m2[(size_t)node * D + d], cnt, prior, cutoff, r, mb);
            RM[(size_t)b * D + d] = make_float2(r, mb);
        }
    }
}

}  // namespace cw

using namespace cw;

extern "C" int64_t cw_rescore_smem_bytes(int32_t D, int32_t max_len, int32_t kc) {
    const int64_t E = (int64_t)kc * max_len;
    return (int64_t)max_len * 8 + (int64_t)RS_WARPS * 32 * 33 * 8 + (int64_t)D * 4 + E * (4 + 4 + 4) + (int64_t)kc * 16 +
           RS_WARPS * 4 + 16 + E * 4 + 16;
}

extern "C" int cw_rescore_rows_build(const cw_store *s, const int32_t *order, int32_t nn, float *rows, void *stream) {
    if (!s || !order || !rows || nn < 1) {
        cw_set_error("cw_rescore_rows_build: bad argument");
        return CW_E_ARG;
    }
    rescore_rows_kernel<<<nn < 148 * 16 ? nn : 148 * 16, 256, 0, (cudaStream_t)stream>>>(*s, order, nn,
                                                                                        reinterpret_cast<float2 *>(rows));
    return cw_check_cuda(cudaGetLastError(), "cw_rescore_rows_build");
}

extern "C" int cw_dense_rescore(const cw_store *s, const cw_index *ix, const float *rows, const int32_t *pos_of_sid,
                                const float *Q, int64_t nq, int kc, const int32_t *cand_sid, const float *cand_score, int k,
                                float hmax, float lmax, float wfac, float eps_scale, int32_t *out_sid, float *out_score,
                                int32_t *fail, void *stream) {
    if (!s || !ix || !rows || !pos_of_sid || !Q || !cand_sid || !cand_score || !out_sid || !out_score || !fail || nq < 0 ||
        k < 1 || kc <= k || kc > CW_RESCORE_MAX_KC || ix->D != s->D || ix->max_len < 1 || !ix->path_idx || !ix->pos_rec ||
        !ix->level_w || (int64_t)kc * ix->max_len > 65535) {
        cw_set_error("cw_dense_rescore: bad argument (k=%d kc=%d max %d, kc*max_len must be <= 65535)", k, kc, CW_RESCORE_MAX_KC);
        return CW_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = cw_check_cuda(cudaMemsetAsync(fail, 0, sizeof(int32_t), st), "cw_dense_rescore: memset");
    if (rc || nq == 0) return rc;
    const int64_t smem = cw_rescore_smem_bytes(s->D, ix->max_len, kc);
    if (smem > 200 * 1024) {
        cw_set_error("cw_dense_rescore: needs %lld bytes of shared memory (D=%d max_len=%d kc=%d)", (long long)smem, s->D,
                     ix->max_len, kc);
        return CW_E_ARG;
    }
    rc = cw_check_cuda(cudaFuncSetAttribute(rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cw_dense_rescore: smem attribute");
    if (rc) return rc;
    RescoreArgs a;
    a.s = *s;
    a.ix = *ix;
    a.RM = reinterpret_cast<const float2 *>(rows);
    a.pos_of_sid = pos_of_sid;
    a.Q = Q;
    a.nq = nq;
    a.kc = kc;
    a.k = k;
    a.cand_sid = cand_sid;
    a.cand_score = cand_score;
    a.inv_prior = 1.0f / s->prior_var;
    a.hmax = hmax;
    a.lmax = lmax;
    a.wfac = wfac;
    a.eps_scale = eps_scale;
    a.mag_scale = 1.1920929e-07f;  // 2^-23: one ulp of the score magnitude per rounding step
    a.out_sid = out_sid;
    a.out_score = out_score;
    a.fail = fail;
    const unsigned grid = (unsigned)(nq < 148 * 64 ? nq : 148 * 64);
    rescore_kernel<<<grid, RS_THREADS, (size_t)smem, st>>>(a);
    return cw_check_cuda(cudaGetLastError(), "cw_dense_rescore");
}
