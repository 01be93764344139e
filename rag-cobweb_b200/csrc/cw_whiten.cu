// cw_whiten.cu -- PCA(+ICA) whitening transform applied in front of ifit / predict
// (PCAICAWhiteningModel.transform, src/whitening/pca_ica.py:30-51):
//     y = ((x - mean) @ P^T / sqrt(lambda + eps)) @ U^T
// Two small fp32 GEMMs per batch (D_in x K, then K x K); the fit (sklearn PCA / FastICA) stays
// on the CPU like in the reference.  One generic kernel does both stages:
//     C[q, j] = (sum_d (A[q, d] - sub[d]) * B[j, d]) / div[j]        (sub / div optional)
// 64 x 64 output tile per CTA, 256 threads, 4 x 4 per thread, k-tiles of 16 staged through shared
// memory.  This work is ~0.1 % of a predict step; it exists so that whitened queries never
// round-trip through host numpy.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cobweb_b200.h"

void cw_set_error(const char *fmt, ...);
int cw_check_cuda(cudaError_t e, const char *what);

namespace cw {

constexpr int WT = 64, WK = 16;

__global__ void __launch_bounds__(256)
affine_gemm_kernel(const float *__restrict__ A, long long nq, int din, const float *__restrict__ B, int dout,
                   const float *__restrict__ sub, const float *__restrict__ div, float *__restrict__ C) {
    __shared__ float As[WK][WT + 1], Bs[WK][WT + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long q0 = (long long)blockIdx.x * WT;
    const int j0 = blockIdx.y * WT;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < din; k0 += WK) {
        // each thread stages 4 elements of A and 4 of B: row = tid / 4 + 0..., 16 k values per row
        for (int i = tid; i < WT * WK; i += 256) {
            const int r = i / WK, kk = i % WK, d = k0 + kk;
            const long long q = q0 + r;
            float a = 0.0f, b = 0.0f;
            if (d < din) {
                if (q < nq) a = A[q * din + d] - (sub ? sub[d] : 0.0f);
                if (j0 + r < dout) b = B[(size_t)(j0 + r) * din + d];
            }
            As[kk][r] = a;
            Bs[kk][r] = b;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < WK; kk++) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { av[i] = As[kk][ty * 4 + i]; bv[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const long long q = q0 + ty * 4 + i;
        if (q >= nq) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int col = j0 + tx * 4 + j;
            if (col < dout) C[q * dout + col] = div ? acc[i][j] / div[col] : acc[i][j];
        }
    }
}

}  // namespace cw

extern "C" int cw_whiten(const float *X, int64_t nq, int32_t din, const float *mean, const float *pca, int32_t k,
                         const float *scale, const float *ica, float *tmp, float *Y, void *stream) {
    if (!X || !pca || !Y || nq < 0 || din < 1 || k < 1 || (ica && !tmp)) {
        cw_set_error("cw_whiten: bad argument");
        return CW_E_ARG;
    }
    if (nq == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)((nq + cw::WT - 1) / cw::WT), (unsigned)((k + cw::WT - 1) / cw::WT));
    float *stage1 = ica ? tmp : Y;
    cw::affine_gemm_kernel<<<grid, 256, 0, st>>>(X, nq, din, pca, k, mean, scale, stage1);
    if (ica) cw::affine_gemm_kernel<<<grid, 256, 0, st>>>(stage1, nq, k, ica, k, nullptr, nullptr, Y);
    return cw_check_cuda(cudaGetLastError(), "cw_whiten");
}
