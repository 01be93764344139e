// cw_common.cuh -- device helpers shared by the strict-arithmetic kernels (ifit, categorize,
// index build).  Translation units including this file are compiled with -fmad=false:
// every elementwise operation must be a single IEEE binary32 operation in the order the
// reference's torch expressions evaluate them (DESIGN.md "Arithmetic contract"), so that
// decisions are bit-identical to the CPU oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cobweb_b200.h"

namespace cw {

constexpr int OP_BEST = 0, OP_NEW = 1, OP_MERGE = 2, OP_SPLIT = 3, OP_LEAF = 4, OP_FRINGE = 5;

// scoring modes derived from cw_store.flags (CobwebTorchTree.compute_score, CobwebTorchTree.py:344-364)
constexpr int MODE_KL = 0;     // use_info && use_kl
constexpr int MODE_INFO = 1;   // use_info && !use_kl
constexpr int MODE_GUESS = 2;  // !use_info

__host__ __device__ inline int mode_of(int flags) {
    if (!(flags & CW_USE_INFO)) return MODE_GUESS;
    return (flags & CW_USE_KL) ? MODE_KL : MODE_INFO;
}

__host__ __device__ inline int pow2_ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// Natural log in binary32 from integer + IEEE add/mul/div only, so the CPU oracle can
// evaluate the identical sequence (fdlibm logf scheme, <1 ulp).  Replaces torch.log on fp32
// tensors (CobwebTorchTree.py:350, CobwebTorchNode.py:102).
__device__ __forceinline__ float logf_strict(float x) {
    const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f;
    const float Lg1 = 0.66666662693f, Lg2 = 0.40000972152f, Lg3 = 0.28498786688f, Lg4 = 0.24279078841f;
    uint32_t ix = __float_as_uint(x);
    int k = 0;
    if (ix < 0x00800000u || ix >= 0x7f800000u) {
        if ((ix << 1) == 0) return -__int_as_float(0x7f800000);
        if (ix >> 31) return __int_as_float(0x7fc00000);
        if (ix >= 0x7f800000u) return x;
        x = x * 33554432.0f;
        k = -25;
        ix = __float_as_uint(x);
    }
    ix += 0x3f800000u - 0x3f3504f3u;
    k += (int)(ix >> 23) - 0x7f;
    ix = (ix & 0x007fffffu) + 0x3f3504f3u;
    x = __uint_as_float(ix);
    float f = x - 1.0f;
    float s = f / (2.0f + f);
    float z = s * s;
    float w = z * z;
    float t1 = w * (Lg2 + w * Lg4);
    float t2 = z * (Lg1 + w * Lg3);
    float R = t2 + t1;
    float hfsq = (0.5f * f) * f;
    float dk = (float)k;
    return ((((s * (hfsq + R)) + (dk * ln2_lo)) - hfsq) + f) + (dk * ln2_hi);
}

// CobwebTorchTree.compute_var (CobwebTorchTree.py:336-342)
__device__ __forceinline__ float var_of(float m2, float count, float prior, bool cutoff) {
    float v = m2 / count;
    if (cutoff) return v < prior ? prior : v;
    return v + prior;
}

// Operands of the FP32-pipe dense score, (x - mean)^2 / var = (x*r + mb)^2 (cw_index.cu builds them into the
// index tiles, cw_rescore.cu re-derives them from the store rows; both must round identically)
__device__ __forceinline__ void dense_operands(float mean, float m2, float count, float prior, bool cutoff, float &r,
                                               float &mb) {
    const float var = count > 0.0f ? var_of(m2, count, prior, cutoff) : prior;
    r = 1.0f / sqrtf(var);
    mb = -(mean * r);
}

// per-attribute transform whose differences / values the score sums use: log var for the
// information-theoretic modes, 1/(2 sqrt(pi) sqrt(var)) for the expected-correct-guess mode
__device__ __forceinline__ float tf_of(float v, int mode) {
    if (mode == MODE_GUESS) {
        const float c = 2.0f * 1.7724539041519165f;  // 2 * torch.sqrt(pi_tensor), sqrtf(fp32 pi) = 0x3fe2dfc5
        return 1.0f / (c * sqrtf(v));
    }
    return logf_strict(v);
}

// the two per-attribute terms of compute_score(mu1, var1, mu2, var2)
__device__ __forceinline__ void score_terms(int mode, float mu1, float v1, float tf1, float mu2, float v2,
                                            float tf2, float &a, float &b) {
    if (mode == MODE_GUESS) {
        a = tf1;
        b = tf2;
    } else {
        a = tf2 - tf1;
        if (mode == MODE_KL) {
            float df = mu1 - mu2;
            b = (v1 + df * df) / v2;
        } else {
            b = 0.0f;
        }
    }
}

// combine the two rounded sums into the score
__device__ __forceinline__ float score_from_sums(int mode, float sa, float sb, int D) {
    if (mode == MODE_KL) {
        float s = sa + sb;
        s = s - (float)D;
        return s / 2.0f;
    }
    if (mode == MODE_INFO) return 0.5f * sa;
    return (-sa) + sb;
}

// torch.isclose(a, b) with default rtol=1e-5, atol=1e-8 (CobwebTorchNode.is_exact_match)
__device__ __forceinline__ bool isclose32(float a, float b) {
    if (a == b) return true;
    float err = fabsf(a - b);
    float allowed = 1e-8f + fabsf(1e-5f * b);
    return isfinite(err) && err <= allowed;
}

// ---- canonical reduction ("tensor.sum()"): balanced pairwise tree in binary64 over groups of
// four consecutive attributes.  A row is handled by a team of Gp = pow2_ceil(ceil(D/4))
// consecutive threads, thread t owning group t.  group4() is level 0; team_reduce() the tree.
__device__ __forceinline__ double group4(float a0, float a1, float a2, float a3) {
    return (((double)a0 + (double)a1) + (double)a2) + (double)a3;
}

// Butterfly over the lanes of one warp that belong to the same team (team width tw <= 32 a
// power of two, teams aligned).  Afterwards every lane of the team holds the tree sum.
template <int K>
__device__ __forceinline__ void warp_tree_reduce(double (&v)[K], int tw) {
    for (int off = 1; off < tw; off <<= 1) {
#pragma unroll
        for (int i = 0; i < K; i++) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
    }
}

}  // namespace cw
