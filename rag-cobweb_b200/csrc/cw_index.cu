// cw_index.cu -- dense prediction index (CobwebWrapper.build_prediction_index,
// src/cobweb/CobwebWrapper.py:186-203) in the operand form of the scoring kernel.
// Strict arithmetic (-fmad=false): sumlog must equal the CPU oracle's bit for bit.
#include "cw_common.cuh"
#include "cw_nvtx.h"

namespace cw {

// sumlog[b] = sum_d log var  (canonical pairwise-binary64 sum); one team of Gp threads per row
__global__ void index_sumlog_kernel(cw_store s, const int *__restrict__ order, int nn, float *sumlog) {
    __shared__ double red[32];
    const int D = s.D, G = (D + 3) / 4, Gp = pow2_ceil(G);
    const int T = blockDim.x, tid = threadIdx.x, NT = T / Gp;
    const int team = tid / Gp, lt = tid % Gp, tw = Gp < 32 ? Gp : 32, wpt = Gp / 32;
    const int warp = tid >> 5, lane = tid & 31;
    const bool cutoff = (s.flags & CW_ACUITY_CUTOFF) != 0;
    const float prior = s.prior_var;
    const int rows_per_pass = NT * gridDim.x;
    for (int base = blockIdx.x * NT; base < nn; base += rows_per_pass) {
        const int b = base + team;
        double acc[1] = {0.0};
        if (b < nn && lt < G) {
            const int node = order[b];
            const float cnt = s.count[node];
            float t[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                int ix = 4 * lt + e;
                if (ix < D) {
                    float var = cnt > 0.0f ? var_of(s.m2[(size_t)node * D + ix], cnt, prior, cutoff) : prior;
                    t[e] = logf_strict(var);
                } else {
                    t[e] = 0.0f;
                }
            }
            acc[0] = group4(t[0], t[1], t[2], t[3]);
        }
        warp_tree_reduce<1>(acc, tw);
        float r = (float)acc[0];
        if (wpt > 1) {
            __syncthreads();
            if (lane == 0) red[warp] = acc[0];
            __syncthreads();
            if ((warp % wpt) == 0) {
                double v = lane < wpt ? red[warp + lane] : 0.0;
                for (int off = 1; off < wpt; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                r = (float)v;
            }
        }
        if (lt == 0 && b < nn) sumlog[b] = r;
    }
}

// R / MB tiles: one CTA per (node tile, k tile); 256 threads, 256/CW_TILE_N threads per node
__global__ void __launch_bounds__(256)
index_tiles_kernel(cw_store s, const int *__restrict__ order, int nn, int n_ktiles, float *R, float *MB) {
    __shared__ float tr[CW_TILE_K][CW_TILE_N + 1], tm[CW_TILE_K][CW_TILE_N + 1];
    const int nt = blockIdx.x, kt = blockIdx.y, tid = threadIdx.x;
    const int D = s.D;
    const bool cutoff = (s.flags & CW_ACUITY_CUTOFF) != 0;
    const float prior = s.prior_var;
    constexpr int TPN = 256 / CW_TILE_N;        // threads per node
    constexpr int KPT = CW_TILE_K / TPN;        // attributes per thread
    const int nl = tid / TPN, part = tid % TPN;
    const int b = nt * CW_TILE_N + nl;
    int node = -1;
    float cnt = 0.0f;
    if (b < nn) { node = order[b]; cnt = s.count[node]; }
#pragma unroll
    for (int e = 0; e < KPT; e++) {
        const int kk = part * KPT + e, d = kt * CW_TILE_K + kk;
        float r = 0.0f, mb = 0.0f;
        if (node >= 0 && d < D)
            dense_operands(s.mean[(size_t)node * D + d], s.m2[(size_t)node * D + d], cnt, prior, cutoff, r, mb);
        tr[kk][nl] = r;
        tm[kk][nl] = mb;
    }
    __syncthreads();
    const size_t tile = ((size_t)nt * n_ktiles + kt) * (CW_TILE_K * CW_TILE_N);
    for (int i = tid; i < CW_TILE_K * CW_TILE_N; i += 256) {
        R[tile + i] = tr[i / CW_TILE_N][i % CW_TILE_N];
        MB[tile + i] = tm[i / CW_TILE_N][i % CW_TILE_N];
    }
}

// RM[b, d] = {r, mb} of index row b (node order[b]), row-major
__global__ void __launch_bounds__(256)
index_rows_kernel(cw_store s, const int *__restrict__ order, int nn, float2 *RM) {
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

void cw_set_error(const char *fmt, ...);
int cw_check_cuda(cudaError_t e, const char *what);

extern "C" int cw_index_build(const cw_store *s, const int32_t *order, int32_t nn, const cw_index *ix, void *stream) {
    CwRange range("cw_index_build");
    if (!s || !order || !ix || nn < 1 || ix->D != s->D || ix->nn != nn || !ix->R || !ix->MB || !ix->sumlog ||
        ix->n_ntiles != (nn + CW_TILE_N - 1) / CW_TILE_N || ix->n_ktiles != (s->D + CW_TILE_K - 1) / CW_TILE_K) {
        cw_set_error("cw_index_build: bad argument / inconsistent index header");
        return CW_E_ARG;
    }
    int Gp = cw::pow2_ceil((s->D + 3) / 4);
    int threads = Gp > 256 ? Gp : 256;
    int nt = threads / Gp;
    int grid = (nn + nt - 1) / nt;
    if (grid > 148 * 16) grid = 148 * 16;
    cw::index_sumlog_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(*s, order, nn, ix->sumlog);
    dim3 g2(ix->n_ntiles, ix->n_ktiles);
    cw::index_tiles_kernel<<<g2, 256, 0, (cudaStream_t)stream>>>(*s, order, nn, ix->n_ktiles, ix->R, ix->MB);
    return cw_check_cuda(cudaGetLastError(), "cw_index_build");
}

extern "C" int cw_index_rows_build(const cw_store *s, const int32_t *order, int32_t nn, float *rows, void *stream) {
    if (!s || !order || !rows || nn < 1) {
        cw_set_error("cw_index_rows_build: bad argument");
        return CW_E_ARG;
    }
    cw::index_rows_kernel<<<nn < 148 * 16 ? nn : 148 * 16, 256, 0, (cudaStream_t)stream>>>(*s, order, nn,
                                                                                          reinterpret_cast<float2 *>(rows));
    return cw_check_cuda(cudaGetLastError(), "cw_index_rows_build");
}
