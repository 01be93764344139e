// cw_tensor.cu -- the dense node score of cobweb_predict_indexed / cobweb_rank_scores
// (src/cobweb/CobwebWrapper.py:232-236, 283-287) as a tensor-core contraction on tcgen05.
//
//   s[q,n] = -0.5*(sumlog[n] + sum_d (x_qd - mu_nd)^2 / var_nd)
//          = h[n] - 0.5 * sum_f A[q,f] * B[n,f]
//   A[q, .] = (x_d^2 , x_d)            B[n, .] = (1/var_nd , -2 mu_nd/var_nd)
//   h[n]    = -0.5*(sumlog[n] + sum_d mu_nd^2/var_nd)          (binary64 at index build)
//
// A single TF32 product keeps 11 significant bits, far too few where a query sits next to a
// leaf (terms of size ~|x|^2/var cancel down to the squared distance).  Every operand is
// therefore split into two TF32 numbers, v = hi + lo (22-23 significant bits), and the product
// is evaluated as hi*hi + hi*lo + lo*hi -- three tcgen05.mma (kind::tf32, fp32 accumulate in
// TMEM) per K step; the dropped lo*lo term is below 2^-22 relative.  Measured against the
// FP32-pipe kernel (cw_dense.cu) the node scores agree to ~1e-6 relative (tests/test_gpu_parity.py).
//
// Kernel shape: persistent, one CTA per SM, 192 threads, CTA tile = 256 queries x 256 nodes:
//   warp 0   producer: cp.async.bulk (TMA bulk copy) of one 64 KB stage = {A_hi, A_lo} 256 queries x 16
//            features + {B_hi, B_lo} 256 nodes x 16 features (8 attributes: 8 "x^2 | 1/var" then 8
//            "x | -2 mean/var"); the operands are stored in HBM as the exact shared-memory image
//            (K-major rows of 64 bytes, 64-byte swizzle), so a stage is two contiguous copies and
//            needs no tensor map; 3 stages;
//   warp 1   MMA issuer: one lane issues 12 tcgen05.mma (M=128, N=256, K=8: two query halves x two K
//            steps x three products) per stage and tcgen05.commit's the stage back to the producer /
//            the accumulators to the epilogue;
//   warps 2-5 epilogue: tcgen05.ld of the two 128x256 fp32 accumulators (lane = query, column =
//            node; together all 512 TMEM columns), h[n] - 0.5*acc, written NODE-major (one 128-byte
//            line per warp store).
// Per stage a CTA moves 64 KB for 12 MMAs (1536 tensor cycles), 43 B/clk: the first version (128 x 256
// tile, 96 KB per 12 MMAs) ran into the L2 throughput cap at 77 % tensor utilisation
// (profiles/r01_tc_score_v1_ncu_full.md).  Tiles are enumerated in panels of query tiles (node tile
// outer, query tile inner) so that the panel's query operands stay in L2 while the node operands stream.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/cobweb_b200.h"

void cw_set_error(const char *fmt, ...);
int cw_check_cuda(cudaError_t e, const char *what);

namespace cwt {

constexpr int TQ = CW_TC_TILE_Q;     // queries per CTA tile (two UMMA M = 128 halves)
constexpr int TM = 128;              // UMMA M
constexpr int TN = CW_TC_TILE_N;     // nodes per tile    (UMMA N)
constexpr int SD = CW_TC_SLAB_D;     // attributes per K slab
constexpr int ROWB = 64;             // bytes per operand row in a slab (16 tf32)
constexpr int A_IMG = TQ * ROWB;     // 16 KB
constexpr int B_IMG = TN * ROWB;     // 16 KB
constexpr int A_BYTES = 2 * A_IMG;   // hi + lo
constexpr int B_BYTES = 2 * B_IMG;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // 64 KB
constexpr int NSTAGE = 3;
constexpr int THREADS = 320;      // producer warp, MMA warp, eight epilogue warps (two per TMEM lane quarter)
constexpr int EPI_THREADS = THREADS - 64;
constexpr int REC_BYTES = 2 * TN * 16;  // two tiles of per-row leaf records (LEAF / FILTER epilogues)
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */ + REC_BYTES;
static_assert(SD * 2 * 4 == ROWB, "a slab row is 8 x^2 features + 8 x features");
static_assert(TQ == 2 * TM, "two accumulators per CTA tile");
static_assert(THREADS - 64 == TN, "one leaf record per epilogue thread");

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol error traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 28)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// ---- tcgen05 wrappers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, one 128 x 256 x 8 TF32 MMA
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp): K-major operand,
// rows of 64 bytes, 64-byte swizzle (Swizzle<2,4,3>); 8-row groups 512 bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address, 16-byte units
    d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(8 * ROWB >> 4) << 32;      // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
    d |= (uint64_t)4 << 61;                    // SWIZZLE_64B
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

#define CW_TMEM_LD32(taddr, v)                                                                                     \
    asm volatile(                                                                                                  \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                  \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                  \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                  \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),          \
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),    \
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),  \
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])   \
        : "r"(taddr)                                                                                               \
        : "memory")

// v = hi + lo with hi, lo representable in TF32 (round to nearest; v - hi is exact in binary32)
__device__ __forceinline__ void split_tf32(float v, float &hi, float &lo) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    hi = __uint_as_float(h);
    const float rem = v - hi;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(rem));
    lo = __uint_as_float(l);
}

// byte offset of 16-byte chunk c (0..3) of row r inside a 64-byte-swizzled operand image: address bits [7,9)
// (row / 2 within the 8-row group) are XORed into the chunk bits [4,6)
__device__ __forceinline__ int swz_off(int r, int c) { return r * ROWB + ((c ^ ((r >> 1) & 3)) << 4); }

// ------------------------------------------------------------------ the scoring kernel
// tile t -> (node tile, query tile): panels of pq query tiles, node tile outer / query tile inner within a panel
__device__ __forceinline__ void tile_coords(long long t, int n_qtiles, int n_ntiles, int pq, int &nt, int &qt) {
    const long long per_full = (long long)pq * n_ntiles;
    const int n_panels = (n_qtiles + pq - 1) / pq;
    int p = (int)(t / per_full);
    if (p > n_panels - 1) p = n_panels - 1;
    const long long rem = t - (long long)p * per_full;
    const int w = min(pq, n_qtiles - p * pq);
    nt = (int)(rem / w);
    qt = p * pq + (int)(rem % w);
}

// What the epilogue does with a finished 128 x 256 accumulator (lane = query, column = index row):
//   EPI_NODE    node score s = h - 0.5 acc, written node-major                              (all rows of an index)
//   EPI_LEAF    leaf score (C[parent] + w s) / len from the cumulative ancestor sums C,
//               written row-major                                           (the sampled leaf tiles of "tf32x3f")
//   EPI_FILTER  the same leaf score, appended to the query's candidate buffer when it reaches the query's
//               threshold tau; nothing else is written                          (the other leaf tiles of "tf32x3f")
enum { EPI_NODE = 0, EPI_LEAF = 1, EPI_FILTER = 2 };
struct TcEpi {
    float *out;              // NODE / LEAF: [rows, ldq]
    long long ldq;
    const float *hconst;     // NODE: per row
    const float4 *leaf_rec;  // LEAF / FILTER: per leaf row {h, w_leaf, 1/len, parent internal row as int bits}
    const float *C;          // LEAF / FILTER: [internal rows, ldq]
    int n_rows;              // FILTER: real rows of this index (the rest is tile padding)
    long long nq;            // FILTER
    const float *tau;        // FILTER: [nq]
    int cap;                 // FILTER: candidate slots per query
    int *cnt;                // FILTER: [nq] candidates appended (may exceed cap: overflow)
    float *cand_val;         // FILTER: [nq, cap]
    int *cand_row;           // FILTER: [nq, cap]
};

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1)
tc_score_kernel(const unsigned char *__restrict__ A, const unsigned char *__restrict__ B, const TcEpi epi, int n_qtiles,
                int nt_begin, int n_ntiles, int n_slabs, int pq) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle atoms need their natural alignment
    const uint32_t bars = base + NSTAGE * STAGE_BYTES;
    // barrier words: full[NSTAGE], empty[NSTAGE], acc_full, acc_empty, then the TMEM base address
    const uint32_t full0 = bars, empty0 = bars + 8 * NSTAGE, accf = bars + 16 * NSTAGE, acce = accf + 8;
    const uint32_t tmem_slot = acce + 8;
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    float4 *recs_s = reinterpret_cast<float4 *>(smem_raw + (bars + 256 - smem_u32(smem_raw)));  // [2][TN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; s++) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(accf, 1);
        mbar_init(acce, EPI_THREADS / 32);  // one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: all 512 columns (two 128 x 256 fp32 accumulators); this warp also frees them
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const long long n_tiles = (long long)n_qtiles * n_ntiles;

    if (warp == 0) {
        // ===== producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                int nt, qt;
                tile_coords(t, n_qtiles, n_ntiles, pq, nt, qt);
                const unsigned char *asrc = A + (size_t)qt * n_slabs * A_BYTES;
                const unsigned char *bsrc = B + (size_t)(nt_begin + nt) * n_slabs * B_BYTES;
                for (int s = 0; s < n_slabs; s++) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t sb = base + stage * STAGE_BYTES;
                    mbar_arrive_expect_tx(full0 + 8 * stage, STAGE_BYTES);
                    bulk_g2s(sb, asrc + (size_t)s * A_BYTES, A_BYTES, full0 + 8 * stage);
                    bulk_g2s(sb + A_BYTES, bsrc + (size_t)s * B_BYTES, B_BYTES, full0 + 8 * stage);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(TM, TN);
            int stage = 0;
            uint32_t phase = 0, aphase = 0;
            for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                mbar_wait(acce, aphase ^ 1);  // epilogue has drained the accumulators of the previous tile
                tc_fence_after();
                for (int s = 0; s < n_slabs; s++) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sb = base + stage * STAGE_BYTES;
                    const uint64_t b_hi = make_smem_desc(sb + A_BYTES), b_lo = make_smem_desc(sb + A_BYTES + B_IMG);
#pragma unroll
                    for (int qh = 0; qh < 2; qh++) {  // the two 128-query halves of the tile, one accumulator each
                        const uint32_t d = tmem_base + (uint32_t)(qh * TN);
                        const uint64_t a_hi = make_smem_desc(sb + qh * (TM * ROWB));
                        const uint64_t a_lo = make_smem_desc(sb + A_IMG + qh * (TM * ROWB));
#pragma unroll
                        for (int k = 0; k < ROWB / 32; k++) {  // 8 TF32 = 32 bytes per MMA; +2 in 16-byte address units
                            const uint64_t ko = (uint64_t)(2 * k);
                            tc_mma_tf32(d, a_hi + ko, b_hi + ko, idesc, (s | k) != 0);
                            tc_mma_tf32(d, a_hi + ko, b_lo + ko, idesc, 1);
                            tc_mma_tf32(d, a_lo + ko, b_hi + ko, idesc, 1);
                        }
                    }
                    tc_commit(empty0 + 8 * stage);  // stage free once these MMAs have read it
                    if (s == n_slabs - 1) tc_commit(accf);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
                aphase ^= 1;
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: warp w reads TMEM lanes 32*(w%4) .. +31 (lane = query of the 128-query half); the two
        // warps of a quarter take alternate 32-column batches
        const int quarter = warp & 3, sub = (warp - 2) >> 2;
        uint32_t aphase = 0;
        int rbuf = 0;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            int nt, qt;
            tile_coords(t, n_qtiles, n_ntiles, pq, nt, qt);
            const long long n0 = (long long)(nt_begin + nt) * TN;
            const float4 *recs = recs_s + rbuf * TN;
            if (MODE != EPI_NODE) {
                // the tile's leaf records go to shared memory while the MMAs of the tile are still running; two
                // buffers, so that one named barrier per tile also protects the buffer of the tile before
                const int e = threadIdx.x - 64;  // 0..255 over the eight epilogue warps
                recs_s[rbuf * TN + e] = __ldg(epi.leaf_rec + n0 + e);
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                rbuf ^= 1;
            }
            const long long ldq = epi.ldq;
            if (MODE == EPI_NODE) {
                mbar_wait(accf, aphase);
                tc_fence_after();
#pragma unroll 1
                for (int qh = 0; qh < 2; qh++) {
                    const long long q0 = (long long)qt * TQ + qh * TM;
                    if (q0 >= ldq) break;  // a half-tile of pure padding past the score matrix
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(qh * TN);
                    const long long q = q0 + quarter * 32 + lane;
#pragma unroll 1
                    for (int c = sub; c < TN / 32; c += 2) {
                        uint32_t v[32];
                        CW_TMEM_LD32(taddr + (uint32_t)(c * 32), v);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const long long n = n0 + c * 32 + j;
                            epi.out[n * ldq + q] = fmaf(-0.5f, __uint_as_float(v[j]), __ldg(epi.hconst + n));
                        }
                    }
                }
            } else {
                // This warp's eight batches of 32 rows: b -> (query half b / 4, column block sub + 2 (b % 4)).  The
                // ancestor sums C[parent][q] of a batch do not depend on the accumulator, so they are fetched one
                // batch ahead -- the first batch while the MMAs of the tile are still running.
                const long long qbase = (long long)qt * TQ + quarter * 32 + lane;
                auto fetch_cp = [&](int b, float (&cp)[32]) {
                    const int qh = b >> 2, c = sub + 2 * (b & 3);
                    const bool qok = (long long)qt * TQ + qh * TM < ldq;
                    const float *Cq = epi.C + qbase + qh * TM;
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const int par = __float_as_int(recs[c * 32 + j].w);  // broadcast read
                        cp[j] = (qok && par >= 0) ? __ldg(Cq + (long long)par * ldq) : 0.0f;
                    }
                };
                float tau0 = 0.0f, tau1 = 0.0f;
                bool live0 = false, live1 = false;
                if (MODE == EPI_FILTER) {
                    live0 = qbase < epi.nq;
                    live1 = qbase + TM < epi.nq;
                    if (live0) tau0 = epi.tau[qbase];
                    if (live1) tau1 = epi.tau[qbase + TM];
                }
                float cpn[32];
                fetch_cp(0, cpn);
                mbar_wait(accf, aphase);
                tc_fence_after();
#pragma unroll 1
                for (int b = 0; b < 8; b++) {
                    const int qh = b >> 2, c = sub + 2 * (b & 3);
                    if ((long long)qt * TQ + qh * TM >= ldq) break;  // a half-tile of pure padding past the score matrix
                    const long long q = qbase + qh * TM;
                    uint32_t v[32];
                    CW_TMEM_LD32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(qh * TN + c * 32), v);
                    float cp[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) cp[j] = cpn[j];
                    if (b + 1 < 8) fetch_cp(b + 1, cpn);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const float tau = qh ? tau1 : tau0;
                    const bool live = qh ? live1 : live0;
                    float sc[32];
                    unsigned hit = 0;
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const float4 r = recs[c * 32 + j];  // broadcast read
                        const float s = fmaf(-0.5f, __uint_as_float(v[j]), r.x);
                        sc[j] = fmaf(r.y, s, cp[j]) * r.z;
                        if (MODE == EPI_LEAF) epi.out[(n0 + c * 32 + j) * ldq + q] = sc[j];
                        else if (sc[j] >= tau) hit |= 1u << j;
                    }
                    if (MODE == EPI_FILTER) {
                        // rows past the index are tile padding (only in the last tile)
                        const long long left = (long long)epi.n_rows - (n0 + c * 32);
                        if (left < 32) hit &= left <= 0 ? 0u : (1u << left) - 1u;
                        if (live && hit) {
                            // one counter update per thread and batch: a returning atomic per hit would put an L2
                            // round trip between the rows
                            int at = atomicAdd(epi.cnt + q, __popc(hit));
#pragma unroll
                            for (int j = 0; j < 32; j++) {
                                if (hit >> j & 1) {
                                    if (at < epi.cap) {
                                        epi.cand_val[q * epi.cap + at] = sc[j];
                                        epi.cand_row[q * epi.cap + at] = (int)(n0 + c * 32 + j);
                                    }
                                    at++;
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acce);
            aphase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// ------------------------------------------------------------------ CTA-pair variant (experimental, COBWEB_B200_TC2=1)
// The same 256 x 256 tile computed by a cluster of two CTAs with tcgen05.mma.cta_group::2: each CTA holds its own 128
// query rows and HALF of the node rows, the accumulator of a CTA is 128 x 256 = 256 TMEM columns, so two tiles fit and
// the epilogue of tile t overlaps the MMAs of tile t+1 (the single-CTA kernel above cannot: its tile fills TMEM).
// Operand traffic per SM is unchanged (32 KB per 6 MMAs).  Node-score epilogue only.
constexpr int P_PART = TM * ROWB;         // 8 KB: 128 rows of one image
constexpr int P_STAGE = 4 * P_PART;       // A_hi, A_lo (own query rows), B_hi, B_lo (own half of the node rows)
constexpr int P_NSTAGE = 6;
constexpr int P_SMEM = P_NSTAGE * P_STAGE + 1024 + 512;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {  // arrives on the barrier at this offset in both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((unsigned short)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
tc_score_pair_kernel(const unsigned char *__restrict__ A, const unsigned char *__restrict__ B, const TcEpi epi, int n_qtiles,
                     int nt_begin, int n_ntiles, int n_slabs, int pq) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + P_NSTAGE * P_STAGE;
    // full[P_NSTAGE], empty[P_NSTAGE], peer_full[P_NSTAGE] (leader), acc_full[2], acc_empty[2] (leader), TMEM base
    const uint32_t full0 = bars, empty0 = full0 + 8 * P_NSTAGE, peer0 = empty0 + 8 * P_NSTAGE, accf0 = peer0 + 8 * P_NSTAGE;
    const uint32_t acce0 = accf0 + 16, tmem_slot = acce0 + 16;
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < P_NSTAGE; s++) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
            mbar_init(peer0 + 8 * s, 1);
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(accf0 + 8 * a, 1);
            mbar_init(acce0 + 8 * a, 2 * (EPI_THREADS / 32));  // every epilogue warp of both CTAs
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // barriers of both CTAs are initialised before anyone signals across
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const long long n_tiles = (long long)n_qtiles * n_ntiles;
    const long long cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == 0) {
        // ===== producer: own query rows + own half of the node rows
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long t = cid; t < n_tiles; t += n_clusters) {
                int nt, qt;
                tile_coords(t, n_qtiles, n_ntiles, pq, nt, qt);
                const unsigned char *asrc = A + (size_t)qt * n_slabs * A_BYTES + rank * P_PART;
                const unsigned char *bsrc = B + (size_t)(nt_begin + nt) * n_slabs * B_BYTES + rank * P_PART;
                for (int s = 0; s < n_slabs; s++) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t sb = base + stage * P_STAGE;
                    mbar_arrive_expect_tx(full0 + 8 * stage, P_STAGE);
                    bulk_g2s(sb, asrc + (size_t)s * A_BYTES, P_PART, full0 + 8 * stage);
                    bulk_g2s(sb + P_PART, asrc + (size_t)s * A_BYTES + A_IMG, P_PART, full0 + 8 * stage);
                    bulk_g2s(sb + 2 * P_PART, bsrc + (size_t)s * B_BYTES, P_PART, full0 + 8 * stage);
                    bulk_g2s(sb + 3 * P_PART, bsrc + (size_t)s * B_BYTES + B_IMG, P_PART, full0 + 8 * stage);
                    if (++stage == P_NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, aphase = 0;
            if (leader) {
                // ===== MMA issuer for the pair
                constexpr uint32_t idesc = make_idesc(2 * TM, TN);
                for (long long t = cid; t < n_tiles; t += n_clusters) {
                    mbar_wait(acce0 + 8 * acc, aphase ^ 1);  // both CTAs have drained this accumulator
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)(acc * TN);
                    for (int s = 0; s < n_slabs; s++) {
                        mbar_wait(full0 + 8 * stage, phase);
                        mbar_wait(peer0 + 8 * stage, phase);
                        tc_fence_after();
                        const uint32_t sb = base + stage * P_STAGE;
                        const uint64_t a_hi = make_smem_desc(sb), a_lo = make_smem_desc(sb + P_PART);
                        const uint64_t b_hi = make_smem_desc(sb + 2 * P_PART), b_lo = make_smem_desc(sb + 3 * P_PART);
#pragma unroll
                        for (int k = 0; k < ROWB / 32; k++) {
                            const uint64_t ko = (uint64_t)(2 * k);
                            tc_mma_tf32_pair(d, a_hi + ko, b_hi + ko, idesc, (s | k) != 0);
                            tc_mma_tf32_pair(d, a_hi + ko, b_lo + ko, idesc, 1);
                            tc_mma_tf32_pair(d, a_lo + ko, b_hi + ko, idesc, 1);
                        }
                        tc_commit_pair(empty0 + 8 * stage);
                        if (s == n_slabs - 1) tc_commit_pair(accf0 + 8 * acc);
                        if (++stage == P_NSTAGE) { stage = 0; phase ^= 1; }
                    }
                    if (++acc == 2) { acc = 0; aphase ^= 1; }
                }
            } else {
                // ===== relay: tell the leader when this CTA's operands of a stage have landed
                const uint32_t peer_remote = mapa_u32(peer0, 0);
                for (long long t = cid; t < n_tiles; t += n_clusters) {
                    for (int s = 0; s < n_slabs; s++) {
                        mbar_wait(full0 + 8 * stage, phase);
                        mbar_arrive_cluster(peer_remote + 8 * stage);
                        if (++stage == P_NSTAGE) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue (both CTAs): this CTA's 128 queries x 256 nodes of the tile
        const int quarter = warp & 3, sub = (warp - 2) >> 2;
        const uint32_t acce_leader = mapa_u32(acce0, 0);
        int acc = 0;
        uint32_t aphase = 0;
        const long long ldq = epi.ldq;
        for (long long t = cid; t < n_tiles; t += n_clusters) {
            int nt, qt;
            tile_coords(t, n_qtiles, n_ntiles, pq, nt, qt);
            const long long n0 = (long long)(nt_begin + nt) * TN;
            mbar_wait(accf0 + 8 * acc, aphase);
            tc_fence_after();
            const long long q0 = (long long)qt * TQ + rank * TM;
            if (q0 < ldq) {
                const long long q = q0 + quarter * 32 + lane;
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * TN);
#pragma unroll 1
                for (int c = sub; c < TN / 32; c += 2) {
                    uint32_t v[32];
                    CW_TMEM_LD32(taddr + (uint32_t)(c * 32), v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const long long n = n0 + c * 32 + j;
                        epi.out[n * ldq + q] = fmaf(-0.5f, __uint_as_float(v[j]), __ldg(epi.hconst + n));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acce_leader + 8 * acc);
            if (++acc == 2) { acc = 0; aphase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer may still be reading this CTA's operands / signalling its barriers
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// ------------------------------------------------------------------ operand builders
// Queries: Q [nq, D] -> A [q tile][slab][hi, lo][256 rows x 64 B swizzled]; row = (x^2 x8 | x x8)
__global__ void __launch_bounds__(256)
tc_queries_kernel(const float *__restrict__ Q, long long nq, int D, int n_slabs, unsigned char *A) {
    const int qt = blockIdx.x, s = blockIdx.y, r = threadIdx.x;
    const long long q = (long long)qt * TQ + r;
    unsigned char *img = A + ((size_t)qt * n_slabs + s) * A_BYTES;
    float x[SD];
#pragma unroll
    for (int e = 0; e < SD; e++) {
        const int d = s * SD + e;
        x[e] = (q < nq && d < D) ? Q[q * D + d] : 0.0f;
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {  // chunks 0,1: squares; chunks 2,3: the values
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float v = x[(c & 1) * 4 + e];
            split_tf32(c < 2 ? v * v : v, hi[e], lo[e]);
        }
        const int off = swz_off(r, c);
        *reinterpret_cast<float4 *>(img + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4 *>(img + A_IMG + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// Nodes: B [node tile][slab][hi, lo][256 rows x 64 B swizzled]; row = (1/var x8 | -2 mean/var x8)
__global__ void __launch_bounds__(256)
tc_nodes_kernel(cw_store s, const int *__restrict__ order, int nn, int n_slabs, unsigned char *B) {
    const int nt = blockIdx.x, sl = blockIdx.y, r = threadIdx.x;
    const int D = s.D;
    const bool cutoff = (s.flags & CW_ACUITY_CUTOFF) != 0;
    const float prior = s.prior_var;
    const int b = nt * TN + r;
    int node = -1;
    float cnt = 0.0f;
    if (b < nn) { node = order[b]; cnt = s.count[node]; }
    unsigned char *img = B + ((size_t)nt * n_slabs + sl) * B_BYTES;
    float iv[SD], mv[SD];
#pragma unroll
    for (int e = 0; e < SD; e++) {
        const int d = sl * SD + e;
        iv[e] = 0.0f;
        mv[e] = 0.0f;
        if (node >= 0 && d < D) {
            float var = prior;
            if (cnt > 0.0f) {
                const float v = s.m2[(size_t)node * D + d] / cnt;  // CobwebTorchTree.compute_var
                var = cutoff ? (v < prior ? prior : v) : v + prior;
            }
            const double inv = 1.0 / (double)var;
            iv[e] = (float)inv;
            mv[e] = (float)(-2.0 * (double)s.mean[(size_t)node * D + d] * inv);
        }
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; e++) split_tf32(c < 2 ? iv[(c & 1) * 4 + e] : mv[(c & 1) * 4 + e], hi[e], lo[e]);
        const int off = swz_off(r, c);
        *reinterpret_cast<float4 *>(img + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4 *>(img + B_IMG + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// h[b] = -0.5*(sumlog[b] + sum_d mean^2/var) in binary64, one warp per index row; padding rows get 0.
__global__ void __launch_bounds__(256)
tc_hconst_kernel(cw_store s, const int *__restrict__ order, int nn, int n_rows, const float *__restrict__ sumlog,
                 float *hconst) {
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= n_rows) return;
    if (b >= nn) {
        if (lane == 0) hconst[b] = 0.0f;
        return;
    }
    const int D = s.D, node = order[b];
    const bool cutoff = (s.flags & CW_ACUITY_CUTOFF) != 0;
    const float prior = s.prior_var, cnt = s.count[node];
    double acc = 0.0;
    for (int d = lane; d < D; d += 32) {
        float var = prior;
        if (cnt > 0.0f) {
            const float v = s.m2[(size_t)node * D + d] / cnt;
            var = cutoff ? (v < prior ? prior : v) : v + prior;
        }
        const double mu = (double)s.mean[(size_t)node * D + d];
        acc += mu * mu / (double)var;
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) hconst[b] = (float)(-0.5 * ((double)sumlog[b] + acc));
}

}  // namespace cwt

using namespace cwt;

extern "C" int64_t cw_tc_a_bytes(int64_t nq, int32_t D) {
    return ((nq + TQ - 1) / TQ) * (int64_t)((D + SD - 1) / SD) * A_BYTES;
}
extern "C" int64_t cw_tc_b_bytes(int32_t nn, int32_t D) {
    return (int64_t)((nn + TN - 1) / TN) * (int64_t)((D + SD - 1) / SD) * B_BYTES;
}

extern "C" int cw_tc_index_build(const cw_store *s, const int32_t *order, int32_t nn, const float *sumlog,
                                 const cw_tc_index *tx, void *stream) {
    if (!s || !order || !sumlog || !tx || nn < 1 || tx->D != s->D || tx->nn != nn || !tx->B || !tx->hconst ||
        tx->n_ntiles != (nn + TN - 1) / TN || tx->n_slabs != (s->D + SD - 1) / SD) {
        cw_set_error("cw_tc_index_build: bad argument / inconsistent header");
        return CW_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    tc_nodes_kernel<<<dim3(tx->n_ntiles, tx->n_slabs), 256, 0, st>>>(*s, order, nn, tx->n_slabs,
                                                                    reinterpret_cast<unsigned char *>(tx->B));
    const int n_rows = tx->n_ntiles * TN;
    tc_hconst_kernel<<<(n_rows + 7) / 8, 256, 0, st>>>(*s, order, nn, n_rows, sumlog, tx->hconst);
    return cw_check_cuda(cudaGetLastError(), "cw_tc_index_build");
}

// C[n][q] = C[parent(n)][q] + w_n * S[n][q] for the internal rows of one tree level, in place on S
__global__ void __launch_bounds__(256)
tc_cumsum_kernel(float *S, long long ldq, int row_begin, int row_end, const int *__restrict__ int_parent,
                 const float *__restrict__ int_w) {
    const long long per_row = ldq >> 2;
    const long long total = (long long)(row_end - row_begin) * per_row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = row_begin + (int)(i / per_row);
        const long long q4 = (i % per_row) << 2;
        const int par = int_parent[n];
        const float w = int_w[n];
        float4 s = *reinterpret_cast<float4 *>(S + (long long)n * ldq + q4);
        float4 c = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (par >= 0) c = *reinterpret_cast<const float4 *>(S + (long long)par * ldq + q4);
        s.x = fmaf(w, s.x, c.x); s.y = fmaf(w, s.y, c.y); s.z = fmaf(w, s.z, c.z); s.w = fmaf(w, s.w, c.w);
        *reinterpret_cast<float4 *>(S + (long long)n * ldq + q4) = s;
    }
}

// Final candidate list of a query: the top-kc of the sampled leaves (sentence ids) united with the leaves the
// filter appended (expanded to their sentences), best kc by (score desc, sentence id asc).  One CTA per query.
constexpr int SEL_MAX = 2048;
__global__ void __launch_bounds__(128)
tc_select_kernel(long long nq, int kc, const int *__restrict__ samp_sid, const float *__restrict__ samp_val, int cap,
                 const int *__restrict__ cnt, const float *__restrict__ cand_val, const int *__restrict__ cand_row,
                 const int *__restrict__ sent_off, const int *__restrict__ sent_ids, int *out_sid, float *out_val, int *ovf) {
    __shared__ float vals[SEL_MAX];
    __shared__ int sids[SEL_MAX];
    __shared__ int total;
    const int tid = threadIdx.x;
    for (long long q = blockIdx.x; q < nq; q += gridDim.x) {
        __syncthreads();
        if (tid == 0) total = 0;
        __syncthreads();
        if (tid < kc && samp_sid && samp_sid[q * kc + tid] >= 0) {
            const int at = atomicAdd(&total, 1);
            vals[at] = samp_val[q * kc + tid];
            sids[at] = samp_sid[q * kc + tid];
        }
        const int n_app = min(cnt[q], cap);
        bool over = cnt[q] > cap;
        for (int i = tid; i < n_app; i += blockDim.x) {
            const int row = cand_row[q * cap + i];
            const float v = cand_val[q * cap + i];
            const int s0 = sent_off[row], s1 = sent_off[row + 1];
            const int at = atomicAdd(&total, s1 - s0);
            for (int s = s0; s < s1; s++)
                if (at + (s - s0) < SEL_MAX) { vals[at + (s - s0)] = v; sids[at + (s - s0)] = sent_ids[s]; }
        }
        __syncthreads();
        const int n = min(total, SEL_MAX);
        over = over || total > SEL_MAX;
        if (tid == 0) ovf[q] = over ? 1 : 0;
        for (int i = tid; i < n; i += blockDim.x) {
            const float v = vals[i];
            const int sid = sids[i];
            int rank = 0;
            for (int j = 0; j < n && rank < kc; j++) rank += vals[j] > v || (vals[j] == v && sids[j] < sid);
            if (rank < kc) { out_sid[q * kc + rank] = sid; out_val[q * kc + rank] = v; }
        }
        for (int r = n + tid; r < kc; r += blockDim.x) { out_sid[q * kc + r] = -1; out_val[q * kc + r] = -__int_as_float(0x7f800000); }
    }
}

static int tc_launch(int mode, const cw_tc_index *tx, const void *a_scratch, int64_t nq, int nt_begin, int nt_count,
                     const TcEpi &epi, cudaStream_t st) {
    const int n_qtiles = (int)((nq + TQ - 1) / TQ);
    auto kern = mode == EPI_NODE ? tc_score_kernel<EPI_NODE> : (mode == EPI_LEAF ? tc_score_kernel<EPI_LEAF> : tc_score_kernel<EPI_FILTER>);
    int rc = cw_check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES),
                           "cw_tc: smem attribute");
    if (rc) return rc;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long n_tiles = (long long)n_qtiles * nt_count;
    if (n_tiles == 0) return 0;
    int grid = (int)(n_tiles < sms ? n_tiles : sms);
    static const bool use_pair = getenv("COBWEB_B200_TC2") && atoi(getenv("COBWEB_B200_TC2")) != 0;
    // query-tile panels: the panel's query operands (A_BYTES per tile and slab) should sit in L2 (~40 MB of it)
    // while the node operands stream through; equal-width panels
    const long long a_tile = (long long)tx->n_slabs * A_BYTES;
    int pq_max = (int)((40ll << 20) / a_tile);
    if (pq_max < 1) pq_max = 1;
    const int n_panels = (n_qtiles + pq_max - 1) / pq_max;
    const int pq = (n_qtiles + n_panels - 1) / n_panels;
    if (use_pair && mode == EPI_NODE) {
        rc = cw_check_cuda(cudaFuncSetAttribute(tc_score_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM),
                           "cw_tc: smem attribute (pair)");
        if (rc) return rc;
        grid = (int)(2 * n_tiles < sms ? 2 * n_tiles : (sms & ~1));
        tc_score_pair_kernel<<<grid, THREADS, P_SMEM, st>>>(reinterpret_cast<const unsigned char *>(a_scratch),
                                                           reinterpret_cast<const unsigned char *>(tx->B), epi, n_qtiles,
                                                           nt_begin, nt_count, tx->n_slabs, pq);
        return cw_check_cuda(cudaGetLastError(), "cw_tc: pair score kernel");
    }
    kern<<<grid, THREADS, SMEM_BYTES, st>>>(reinterpret_cast<const unsigned char *>(a_scratch),
                                            reinterpret_cast<const unsigned char *>(tx->B), epi, n_qtiles, nt_begin, nt_count,
                                            tx->n_slabs, pq);
    return cw_check_cuda(cudaGetLastError(), "cw_tc: score kernel");
}

extern "C" int cw_tc_build_queries(const cw_tc_index *tx, const float *Q, int64_t nq, void *a_scratch, void *stream) {
    if (!tx || !Q || !a_scratch || nq < 0) {
        cw_set_error("cw_tc_build_queries: bad argument");
        return CW_E_ARG;
    }
    if (nq == 0) return 0;
    const int n_qtiles = (int)((nq + TQ - 1) / TQ);
    tc_queries_kernel<<<dim3(n_qtiles, tx->n_slabs), 256, 0, (cudaStream_t)stream>>>(Q, nq, tx->D, tx->n_slabs,
                                                                                    reinterpret_cast<unsigned char *>(a_scratch));
    return cw_check_cuda(cudaGetLastError(), "cw_tc_build_queries");
}

extern "C" int cw_tc_score_tiles(const cw_tc_index *tx, const void *a_scratch, int64_t nq, int mode, int32_t nt_begin,
                                 int32_t nt_count, float *out, int64_t ldq, const float *C, const float *leaf_rec,
                                 int32_t n_rows, const float *tau, int32_t cap, int32_t *cnt, float *cand_val,
                                 int32_t *cand_row, void *stream) {
    if (!tx || !a_scratch || nq < 0 || ldq < cw_score_ldq(nq) || (ldq % TM) || mode < EPI_NODE || mode > EPI_FILTER ||
        nt_begin < 0 || nt_count < 0 || nt_begin + nt_count > tx->n_ntiles || (mode != EPI_FILTER && !out) ||
        (mode != EPI_NODE && (!C || !leaf_rec)) || (mode == EPI_FILTER && (!tau || !cnt || !cand_val || !cand_row || cap < 1))) {
        cw_set_error("cw_tc_score_tiles: bad argument (mode %d, tiles %d+%d of %d)", mode, nt_begin, nt_count, tx ? tx->n_ntiles : -1);
        return CW_E_ARG;
    }
    if (nq == 0) return 0;
    TcEpi epi;
    epi.out = out;
    epi.ldq = ldq;
    epi.hconst = tx->hconst;
    epi.leaf_rec = reinterpret_cast<const float4 *>(leaf_rec);
    epi.C = C;
    epi.n_rows = n_rows;
    epi.nq = nq;
    epi.tau = tau;
    epi.cap = cap;
    epi.cnt = cnt;
    epi.cand_val = cand_val;
    epi.cand_row = cand_row;
    return tc_launch(mode, tx, a_scratch, nq, nt_begin, nt_count, epi, (cudaStream_t)stream);
}

extern "C" int cw_tc_cumsum_level(float *S, int64_t ldq, int32_t row_begin, int32_t row_end, const int32_t *int_parent,
                                  const float *int_w, void *stream) {
    if (!S || !int_parent || !int_w || (ldq & 3) || row_begin < 0 || row_end < row_begin) {
        cw_set_error("cw_tc_cumsum_level: bad argument");
        return CW_E_ARG;
    }
    if (row_end == row_begin) return 0;
    const long long total = (long long)(row_end - row_begin) * (ldq >> 2);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    tc_cumsum_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(S, ldq, row_begin, row_end, int_parent, int_w);
    return cw_check_cuda(cudaGetLastError(), "cw_tc_cumsum_level");
}

extern "C" int cw_tc_select(int64_t nq, int kc, const int32_t *samp_sid, const float *samp_val, int32_t cap,
                            const int32_t *cnt, const float *cand_val, const int32_t *cand_row, const int32_t *sent_off,
                            const int32_t *sent_ids, int32_t *out_sid, float *out_val, int32_t *ovf, void *stream) {
    if (nq < 0 || kc < 1 || kc > 128 || cap < 1 || !cnt || !cand_val || !cand_row || !sent_off || !sent_ids || !out_sid ||
        !out_val || !ovf || (samp_sid && !samp_val)) {
        cw_set_error("cw_tc_select: bad argument");
        return CW_E_ARG;
    }
    if (nq == 0) return 0;
    tc_select_kernel<<<(unsigned)(nq < 148 * 16 ? nq : 148 * 16), 128, 0, (cudaStream_t)stream>>>(
        nq, kc, samp_sid, samp_val, cap, cnt, cand_val, cand_row, sent_off, sent_ids, out_sid, out_val, ovf);
    return cw_check_cuda(cudaGetLastError(), "cw_tc_select");
}

extern "C" int cw_dense_node_scores_tc(const cw_tc_index *tx, const float *Q, int64_t nq, void *a_scratch,
                                       float *node_scores, int64_t ldq, void *stream) {
    if (!tx || !Q || !a_scratch || !node_scores || nq < 0 || ldq < cw_score_ldq(nq) || (ldq % TM)) {
        cw_set_error("cw_dense_node_scores_tc: bad argument (ldq must be >= cw_score_ldq(nq) and a multiple of %d)", TM);
        return CW_E_ARG;
    }
    if (nq == 0) return 0;
    int rc = cw_tc_build_queries(tx, Q, nq, a_scratch, stream);
    if (rc) return rc;
    return cw_tc_score_tiles(tx, a_scratch, nq, EPI_NODE, 0, tx->n_ntiles, node_scores, ldq, nullptr, nullptr, 0, nullptr, 0,
                             nullptr, nullptr, nullptr, stream);
}
