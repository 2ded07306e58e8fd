// Matching stage beyond 1024 keypoints (BASELINE config 5: 1080p, K = 2048): the kernel matrix K = exp(-cost/eps) of a
// descriptor pair (16.8 MB at 2048 x 2048) fits no cluster, so it lives in global memory in a padded, 16-byte aligned
// layout and every Sinkhorn iteration is ONE streaming sweep of it.
//
// Replaces matching/sinkhorn.py:79-208 in the scaling form (a = mu / (K b), b = nu / (K^T a), P = a K b), like the hybrid
// cluster kernel (sinkhorn_hy.cu) and the older generic kernels (sinkhorn.cu: xd_row_kernel / cost_tc_kernel), which these
// kernels supersede wherever they are eligible:
//
//   cost_pk_kernel      similarity GEMM on tcgen05 from the operands packed once by pack_f16_kernel (two fp16 terms,
//                       hi*hi + hi*lo + lo*hi, FP32 accumulators): persistent CTAs, 128 (columns j) x 256 (rows i) tiles,
//                       operand chunks by 1-D bulk copies through a 4-stage ring, accumulators double-buffered in tensor
//                       memory so that the exp epilogue of tile t runs under the MMAs of tile t+1.
//                       (cost_tc_kernel: 64 x 128 tiles, operands converted by producer warps in every CTA, no overlap.)
//   xs_sweep_kernel     one iteration = one pass over K: persistent CTAs, 8-row stages fetched by bulk copies into a
//                       3-slot ring; a stage's rows give their row sums (warp per row) -> a_i, then every thread adds
//                       a_i K_ij into the running sums of its own columns; a unit of 16 rows ends in one partial row.
//                       Odd iterations walk the units backwards: what the previous sweep read last is still in L2.
//   xs_col_kernel       b_j = nu_j / (sum of the unit partials), fixed order.
//   xs_finalize_kernel  P = a K b, rows staged through shared memory to coalesced stores into the (N+1) x (M+1) output.
//
// A unit is a function of N alone and partials are combined in a fixed order, so results do not depend on the batch size,
// on the number of CTAs or on which CTA processed which unit.
#include <cuda_fp16.h>
#include <math_constants.h>

#include "common.cuh"

namespace om {

// sinkhorn_hy.cu: descriptor rows -> two fp16 terms in the MMA's core-matrix layout [z][tile][chunk][term][g][Rp][8 halves]
int pack_f16_launch(const float* d, int B, int rows, int D, int Rp, int tiles, int G, unsigned char* out, float* norms,
                    unsigned int* ovf, cudaStream_t st);

namespace {

__device__ __forceinline__ float ex2x(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load_hint(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     dst_smem), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
// L2 policy of the matrix reads: a fixed (address-hashed) fraction of the lines is kept with priority, the rest streams
// through -- with a matrix larger than L2 the kept part survives from one sweep to the next whatever the order
__device__ __forceinline__ uint64_t l2_policy(int id) {
    uint64_t pol = 0;
    switch (id) {
        case 1: asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol)); break;
        case 2: asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 0.5;" : "=l"(pol)); break;
        case 3: asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 0.75;" : "=l"(pol)); break;
        case 4: asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 0.25;" : "=l"(pol)); break;
        case 5: asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol)); break;
        case 6: asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 0.625;" : "=l"(pol)); break;
        default: break;
    }
    return pol;
}
// UMMA shared-memory descriptor, no swizzle, K-major (cute/arch/mma_sm100_desc.hpp layout)
__device__ __forceinline__ uint64_t xdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}
__device__ __forceinline__ void xmma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(accumulate) : "memory");
}
__device__ __forceinline__ void xmma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void xtmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// cost_pk_kernel
// ------------------------------------------------------------------------------------------
constexpr int XA = 128;                            // columns j of K (rows of d2) per tile: MMA M, tensor-memory lanes
constexpr int XB = 256;                            // rows i of K (rows of d1) per tile: MMA N, tensor-memory columns
constexpr int XG = 4;                              // 16-byte K groups per chunk (32 descriptor elements)
constexpr int XC_STAGES = 4;
constexpr int XA_TERM = XG * XA * 16;              // bytes of one fp16 term of a d2 chunk (8 KB)
constexpr int XB_TERM = XG * XB * 16;              // ... of a d1 chunk (16 KB)
constexpr int XC_STAGE = 2 * XA_TERM + 2 * XB_TERM;   // 48 KB
constexpr int XC_EPI_WARPS = 16;                   // four per tensor-memory lane quadrant, 64 accumulator columns each
constexpr int XC_THREADS = 64 + 32 * XC_EPI_WARPS;  // warp 0: copy issuer, warp 1: MMA issuer, the others: epilogue
constexpr size_t XC_SMEM = (size_t)XC_STAGES * XC_STAGE + 128;

struct CostArgs {
    const unsigned char* d1p;                      // [z][T1][chunk][term][g][256][8 halves]
    const unsigned char* d2p;                      // [z][T2][chunk][term][g][128][8 halves]
    const float* n1;                               // [z][T1 * 256] squared norms (padding rows: 0)
    const float* n2;                               // [z][T2 * 128]
    const unsigned int* ovf;                       // per pair: descriptors out of fp16 range -> the pair is left to cost_l2_kernel
    float* Kmat;                                   // [z][Np8][Mp]
    int N, M, D, Mp, T1, T2, total;
    size_t zstride;                                // Np8 * Mp
    float scale2;                                  // log2(e) / eps
};

__global__ void __launch_bounds__(XC_THREADS, 1) cost_pk_kernel(CostArgs a) {
    extern __shared__ __align__(128) unsigned char xsm[];
    // [0..3] stage full, [4..7] stage free, [8,9] accumulator full, [10,11] accumulator free
    uint64_t* bars = reinterpret_cast<uint64_t*>(xsm + (size_t)XC_STAGES * XC_STAGE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < 10; ++i) mbar_init(smem_u32(&bars[i]), 1u);
        mbar_init(smem_u32(&bars[10]), (uint32_t)XC_EPI_WARPS);
        mbar_init(smem_u32(&bars[11]), (uint32_t)XC_EPI_WARPS);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t stage0 = smem_u32(xsm);
    const int nchunks = a.D / (8 * XG);
    const int tpp = a.T1 * a.T2;

    if (warp == 0) {
        if (lane == 0) {
            int k = 0;
            for (int t = blockIdx.x; t < a.total; t += gridDim.x) {
                const int z = t / tpp, r = t - z * tpp, ib = r / a.T2, jb = r - ib * a.T2;
                if (a.ovf[z] != 0u) continue;
                const unsigned char* gA = a.d2p + ((size_t)z * a.T2 + jb) * nchunks * (size_t)(2 * XA_TERM);
                const unsigned char* gB = a.d1p + ((size_t)z * a.T1 + ib) * nchunks * (size_t)(2 * XB_TERM);
                for (int c = 0; c < nchunks; ++c, ++k) {
                    const int s = k % XC_STAGES;
                    if (k >= XC_STAGES) mbar_wait(smem_u32(&bars[4 + s]), (uint32_t)((k / XC_STAGES - 1) & 1));
                    const uint32_t full = smem_u32(&bars[s]);
                    mbar_arrive_expect_tx(full, (uint32_t)XC_STAGE);
                    bulk_load(stage0 + s * XC_STAGE, gA + (size_t)c * (2 * XA_TERM), 2 * XA_TERM, full);
                    bulk_load(stage0 + s * XC_STAGE + 2 * XA_TERM, gB + (size_t)c * (2 * XB_TERM), 2 * XB_TERM, full);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(XB >> 3) << 17) | ((uint32_t)(XA >> 4) << 24);   // D=F32, A=B=F16, K-major
            constexpr uint32_t A_LBO = XA * 16, B_LBO = XB * 16;
            int k = 0, tl = 0;
            for (int t = blockIdx.x; t < a.total; t += gridDim.x) {
                const int z = t / tpp;
                if (a.ovf[z] != 0u) continue;
                const int buf = tl & 1;
                if (tl >= 2) mbar_wait(smem_u32(&bars[10 + buf]), (uint32_t)((tl / 2 - 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + (uint32_t)(buf * XB);
                for (int c = 0; c < nchunks; ++c, ++k) {
                    const int s = k % XC_STAGES;
                    mbar_wait(smem_u32(&bars[s]), (uint32_t)((k / XC_STAGES) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = stage0 + s * XC_STAGE;
#pragma unroll
                    for (int ks = 0; ks < XG / 2; ++ks) {
                        const uint64_t ahi = xdesc(sa + 2 * ks * A_LBO, A_LBO, 128);
                        const uint64_t alo = xdesc(sa + XA_TERM + 2 * ks * A_LBO, A_LBO, 128);
                        const uint64_t bhi = xdesc(sa + 2 * XA_TERM + 2 * ks * B_LBO, B_LBO, 128);
                        const uint64_t blo = xdesc(sa + 2 * XA_TERM + XB_TERM + 2 * ks * B_LBO, B_LBO, 128);
                        xmma_f16(d, ahi, bhi, IDESC, (c | ks) != 0);
                        xmma_f16(d, ahi, blo, IDESC, 1u);
                        xmma_f16(d, alo, bhi, IDESC, 1u);
                    }
                    xmma_commit(smem_u32(&bars[4 + s]));
                }
                xmma_commit(smem_u32(&bars[8 + buf]));
                ++tl;
            }
        }
    } else {
        // epilogue warp (q, h): tensor-memory lanes 32q..32q+31 (the quadrant a warp may read is warp % 4), accumulator columns
        // CW h .. CW h + CW - 1: lane = column j of K, registers = 32 consecutive rows i -> every store instruction writes 128
        // contiguous bytes of one row
        constexpr int CW = XB / (XC_EPI_WARPS / 4);
        const int q = warp & 3, h = (warp - 2) >> 2;
        int tl = 0;
        for (int t = blockIdx.x; t < a.total; t += gridDim.x) {
            const int z = t / tpp, r = t - z * tpp, ib = r / a.T2, jb = r - ib * a.T2;
            if (a.ovf[z] != 0u) continue;
            const int buf = tl & 1;
            mbar_wait(smem_u32(&bars[8 + buf]), (uint32_t)((tl / 2) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int j = jb * XA + 32 * q + lane;
            const bool jok = j < a.M;
            const float n2j = a.n2[((size_t)z * a.T2 + jb) * XA + 32 * q + lane];
            const float* n1t = a.n1 + ((size_t)z * a.T1 + ib) * XB + CW * h;
            float* Kz = a.Kmat + (size_t)z * a.zstride + j;
            const int i_base = ib * XB + CW * h;
#pragma unroll 1
            for (int cc = 0; cc < CW / 32; ++cc) {
                uint32_t acc[32];
                xtmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * XB + CW * h + 32 * cc), acc);
#pragma unroll
                for (int ii = 0; ii < 32; ++ii) {
                    const int i = i_base + 32 * cc + ii;
                    const float dot = __uint_as_float(acc[ii]);
                    const float cost = fmaxf(__fsub_rn(__fadd_rn(__ldg(n1t + 32 * cc + ii), n2j), __fmul_rn(2.0f, dot)), 0.0f);   // sinkhorn.py:98-103
                    if (jok && i < a.N) Kz[(size_t)i * a.Mp] = ex2x(__fmul_rn(-cost, a.scale2));
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[10 + buf]));
            ++tl;
        }
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

// dustbin row N and column M (sinkhorn.py:182-187) = exp(-unused/eps); padding rows (N, Np8) and columns (M, Mp) = 0
__global__ void __launch_bounds__(256) xl_fill_kernel(float* Kmat, int N, int M, int Np8, int Mp, size_t zstride, float kd) {
    float* Kz = Kmat + (size_t)blockIdx.y * zstride;
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e < Mp) {
        Kz[(size_t)N * Mp + e] = e <= M ? kd : 0.0f;
        for (int r = N + 1; r < Np8; ++r) Kz[(size_t)r * Mp + e] = 0.0f;
    }
    if (e < N) {
        Kz[(size_t)e * Mp + M] = kd;
        for (int c = M + 1; c < Mp; ++c) Kz[(size_t)e * Mp + c] = 0.0f;
    }
}

// b = 1 on the real columns and the dustbin column, 0 on the padding
__global__ void __launch_bounds__(256) xl_init_b_kernel(float* b, int M, int Mp) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j < Mp) b[(size_t)blockIdx.y * Mp + j] = j <= M ? 1.0f : 0.0f;
}

// ------------------------------------------------------------------------------------------
// xs_sweep_kernel
// ------------------------------------------------------------------------------------------
constexpr int XS_ROWS = 8;                         // rows per stage == consumer warps
constexpr int XS_SLOTS = 3;
constexpr int XS_UNIT = 2;                         // stages per unit (16 rows -> one partial row)
constexpr int XS_CONS = XS_ROWS * 32;
constexpr int XS_THREADS = XS_CONS + 32;           // + the copy issuer's warp
constexpr int XS_MAX_MP = 2148;                    // 27 rows of Mp floats + barriers within 227 KB
constexpr int XS_NQ = (XS_MAX_MP / 4 + 31) / 32;   // float4 column groups of a row per lane (17)

struct SweepArgs {
    const float* Kmat;
    const float* b;                                // [z][Mp]
    float* a;                                      // [z][Np8]
    float* Tpart;                                  // [z][upp][Mp]
    int N, Np8, Mp, upp, total, reverse;
    int policy;                                    // L2 policy of the matrix reads (l2_policy)
    size_t zstride;
    float mu_dust;                                 // mu_N = M (sinkhorn.py:197-198)
};

__global__ void __launch_bounds__(XS_THREADS, 1) xs_sweep_kernel(SweepArgs p) {
    extern __shared__ __align__(128) float ssm[];
    const int Mp = p.Mp, Mp4 = Mp >> 2;
    float* sK = ssm;                                               // [XS_SLOTS][XS_ROWS][Mp]
    float* sBb = ssm + (size_t)XS_SLOTS * XS_ROWS * Mp;            // [3][Mp]: b of the unit being processed (by unit count % 3)
    float* sA = sBb + (size_t)3 * Mp;                              // [2][XS_ROWS]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + 16);         // [0..2] slot full, [3..5] slot free, [6..8] b landed
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < XS_SLOTS; ++i) {
            mbar_init(smem_u32(&bars[i]), 1u);
            mbar_init(smem_u32(&bars[XS_SLOTS + i]), (uint32_t)XS_ROWS);
            mbar_init(smem_u32(&bars[2 * XS_SLOTS + i]), 1u);
        }
        mbar_fence_init();
    }
    __syncthreads();
    // programmatic dependent launch: the next kernel of the stream may be scheduled now.  The matrix is read-only during the
    // iterations, so its first stages are fetched at once, under the column kernel that is still finishing b; everything that
    // depends on the previous kernels (b) or that they still read (the partial rows) waits for them (griddepcontrol.wait).
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t stage_bytes = (uint32_t)(XS_ROWS * Mp * 4);
    // a CTA owns a contiguous run of units and walks it forwards (even iterations) or backwards (odd): the same SM re-reads
    // what it read last
    const int u_begin = (int)((long long)p.total * blockIdx.x / gridDim.x), u_end = (int)((long long)p.total * (blockIdx.x + 1) / gridDim.x);

    if (warp == XS_ROWS) {
        if (lane == 0) {
            const uint64_t pol = l2_policy(p.policy & 7);
            bool waited = false;
            int k = 0, n = 0;
            for (int u = u_begin; u < u_end; ++u, ++n) {
                const int uu = p.reverse ? u_end - 1 - (u - u_begin) : u;
                const int z = uu / p.upp, q = uu - z * p.upp;
                const int row0 = q * (XS_UNIT * XS_ROWS);
                const int nst = min(XS_UNIT, (p.Np8 - row0) / XS_ROWS);
                const float* Kz = p.Kmat + (size_t)z * p.zstride + (size_t)row0 * Mp;
                for (int s = 0; s < nst; ++s, ++k) {
                    const int slot = k % XS_SLOTS;
                    if (k >= XS_SLOTS) mbar_wait(smem_u32(&bars[XS_SLOTS + slot]), (uint32_t)((k / XS_SLOTS - 1) & 1));
                    const uint32_t full = smem_u32(&bars[slot]);
                    mbar_arrive_expect_tx(full, stage_bytes);
                    if ((p.policy & 7) != 0) {
                        bulk_load_hint(smem_u32(sK + (size_t)slot * XS_ROWS * Mp), Kz + (size_t)s * XS_ROWS * Mp, stage_bytes, full, pol);
                    } else {
                        bulk_load(smem_u32(sK + (size_t)slot * XS_ROWS * Mp), Kz + (size_t)s * XS_ROWS * Mp, stage_bytes, full);
                    }
                    if (s == 0) {
                        // b of the unit's pair (buffer n % 3: free once the unit three back is done, which the slot wait above
                        // implies); the first one of the kernel waits for the column kernel
                        if (!waited) {
                            asm volatile("griddepcontrol.wait;" ::: "memory");
                            waited = true;
                        }
                        const uint32_t bb = smem_u32(&bars[2 * XS_SLOTS + n % 3]);
                        mbar_arrive_expect_tx(bb, (uint32_t)(Mp * 4));
                        bulk_load(smem_u32(sBb + (size_t)(n % 3) * Mp), p.b + (size_t)z * Mp, (uint32_t)(Mp * 4), bb);
                    }
                }
            }
        }
        return;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");

    float4 acc[3];
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) acc[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = 0, n = 0;
    for (int u = u_begin; u < u_end; ++u, ++n) {
        const int uu = p.reverse ? u_end - 1 - (u - u_begin) : u;
        const int z = uu / p.upp, q = uu - z * p.upp;
        const int row0 = q * (XS_UNIT * XS_ROWS);
        const int nst = min(XS_UNIT, (p.Np8 - row0) / XS_ROWS);
        const float4* b4 = reinterpret_cast<const float4*>(sBb + (size_t)(n % 3) * Mp);
        mbar_wait(smem_u32(&bars[2 * XS_SLOTS + n % 3]), (uint32_t)((n / 3) & 1));
        // this lane's columns of b, in registers for the unit's stages (read from shared memory in every row pass they were
        // half of its shared-memory traffic)
        float4 breg[XS_NQ];
#pragma unroll
        for (int i = 0; i < XS_NQ; ++i) {
            const int c = lane + 32 * i;
            breg[i] = c < Mp4 ? b4[c] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int s = 0; s < nst; ++s, ++k) {
            const int slot = k % XS_SLOTS;
            mbar_wait(smem_u32(&bars[slot]), (uint32_t)((k / XS_SLOTS) & 1));
            const float4* st4 = reinterpret_cast<const float4*>(sK + (size_t)slot * XS_ROWS * Mp);
            // ---- a_i = mu_i / sum_j K_ij b_j: one warp per row ----
            {
                const float4* row4 = st4 + (size_t)warp * Mp4;
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int i = 0; i < XS_NQ; ++i) {
                    const int c = lane + 32 * i;
                    if (c < Mp4) {
                        const float4 kv = row4[c];
                        s0 = fmaf(kv.x, breg[i].x, s0); s1 = fmaf(kv.y, breg[i].y, s1);
                        s2 = fmaf(kv.z, breg[i].z, s2); s3 = fmaf(kv.w, breg[i].w, s3);
                    }
                }
                const float rs = warp_sum((s0 + s1) + (s2 + s3));
                const int gi = row0 + s * XS_ROWS + warp;
                const float av = gi <= p.N ? __fdividef(gi == p.N ? p.mu_dust : 1.0f, rs) : 0.0f;
                if (lane == 0) {
                    sA[(k & 1) * XS_ROWS + warp] = av;
                    p.a[(size_t)z * p.Np8 + gi] = av;
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(XS_CONS) : "memory");
            // ---- column sums: every thread adds a_i K_ij of the stage's rows into the sums of its own columns ----
            {
                const float* av8 = sA + (k & 1) * XS_ROWS;
                float ar[XS_ROWS];
#pragma unroll
                for (int r = 0; r < XS_ROWS; ++r) ar[r] = av8[r];
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    const int g = tid + XS_CONS * jj;
                    if (g < Mp4) {
                        float4 t = acc[jj];
#pragma unroll
                        for (int r = 0; r < XS_ROWS; ++r) {
                            const float4 kv = st4[(size_t)r * Mp4 + g];
                            t.x = fmaf(kv.x, ar[r], t.x); t.y = fmaf(kv.y, ar[r], t.y);
                            t.z = fmaf(kv.z, ar[r], t.z); t.w = fmaf(kv.w, ar[r], t.w);
                        }
                        acc[jj] = t;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[XS_SLOTS + slot]));
        }
        float4* tp = reinterpret_cast<float4*>(p.Tpart + ((size_t)z * p.upp + q) * Mp);
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) {
            const int g = tid + XS_CONS * jj;
            if (g < Mp4) tp[g] = acc[jj];
            acc[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// b_j = nu_j / sum over the units' partial rows; 32 columns per CTA, warp g adds the partials g, g + 8, ... (128-byte rows):
// all of a thread's loads are issued before the first addition (one L2 round trip instead of a chain of them), the sums run
// in a fixed order, the eight group sums are combined in a fixed order
constexpr int XCOL_MAXQ = 17;                      // partial rows per warp held in registers: upp <= 136, i.e. N <= 2175
__global__ void __launch_bounds__(256) xs_col_kernel(const float* Tpart, float* b, int upp, int M, int Mp, float nu_dust) {
    __shared__ float sT[8][33];
    const int z = blockIdx.y;
    const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + c;
    float t0 = 0.0f, t1 = 0.0f;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (j <= M) {
        const float* tp = Tpart + (size_t)z * upp * Mp + j;
        if (upp <= 8 * XCOL_MAXQ) {
            float v[XCOL_MAXQ];
#pragma unroll
            for (int k = 0; k < XCOL_MAXQ; ++k) v[k] = g + 8 * k < upp ? __ldcg(tp + (size_t)(g + 8 * k) * Mp) : 0.0f;
#pragma unroll
            for (int k = 0; k < XCOL_MAXQ; ++k) {
                if (k & 1) t1 += v[k];
                else t0 += v[k];
            }
        } else {
            int q = g;
            for (; q + 8 < upp; q += 16) { t0 += tp[(size_t)q * Mp]; t1 += tp[(size_t)(q + 8) * Mp]; }
            if (q < upp) t0 += tp[(size_t)q * Mp];
        }
    }
    sT[g][c] = t0 + t1;
    __syncthreads();
    if (g == 0 && j <= M) {
        float t = sT[0][c];
#pragma unroll
        for (int k = 1; k < 8; ++k) t += sT[k][c];
        b[(size_t)z * Mp + j] = __fdividef(j == M ? nu_dust : 1.0f, t);   // nu_M = N (sinkhorn.py:199-200)
    }
}

// P = a_i K_ij b_j (sinkhorn.py:145, :206): one warp per row, aligned 16-byte reads of K and b, the row staged in shared
// memory and written with coalesced 4-byte stores (rows of the (N+1) x (M+1) output are not 16-byte aligned)
__global__ void __launch_bounds__(256) xs_finalize_kernel(const float* Kmat, const float* a, const float* b, float* P, int N, int M,
                                                          int Np8, int Mp, size_t zstride) {
    extern __shared__ __align__(16) float fsm[];
    const int z = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Mp4 = Mp >> 2;
    float* st = fsm + (size_t)warp * Mp;
    const float4* b4 = reinterpret_cast<const float4*>(b + (size_t)z * Mp);
    for (int i = blockIdx.x * 8 + warp; i <= N; i += gridDim.x * 8) {
        const float ai = a[(size_t)z * Np8 + i];
        const float4* row4 = reinterpret_cast<const float4*>(Kmat + (size_t)z * zstride + (size_t)i * Mp);
#pragma unroll 4
        for (int c = lane; c < Mp4; c += 32) {
            const float4 kv = __ldcs(row4 + c), bv = __ldg(b4 + c);
            reinterpret_cast<float4*>(st)[c] = make_float4((ai * kv.x) * bv.x, (ai * kv.y) * bv.y, (ai * kv.z) * bv.z, (ai * kv.w) * bv.w);
        }
        __syncwarp();
        // 16-byte stores: the row starts `o` floats past a 16-byte boundary of P, so the aligned chunks of the output are
        // chunks of the staged row shifted by j0 = (4 - o) % 4 elements (two aligned shared-memory reads and a select)
        const size_t rowoff = ((size_t)z * (N + 1) + i) * (M + 1);
        float* out = P + rowoff;
        const int j0 = (int)((4 - (rowoff & 3)) & 3);
        const int nch = (M + 1 - j0) >> 2;                           // whole 4-float chunks out[j0 + 4m .. j0 + 4m + 3], all <= M
        if (lane < j0 && lane <= M) __stcs(out + lane, st[lane]);
        const float4* st4 = reinterpret_cast<const float4*>(st);
        for (int m = lane; m < nch; m += 32) {
            const float4 A = st4[m];
            float4 v = A;
            if (j0 != 0) {
                const float4 Bv = st4[m + 1];
                v = j0 == 1 ? make_float4(A.y, A.z, A.w, Bv.x) : (j0 == 2 ? make_float4(A.z, A.w, Bv.x, Bv.y) : make_float4(A.w, Bv.x, Bv.y, Bv.z));
            }
            __stcs(reinterpret_cast<float4*>(out + j0 + 4 * m), v);
        }
        {
            const int j = j0 + 4 * nch + lane;
            if (lane < 4 && j <= M) __stcs(out + j, st[j]);
        }
        __syncwarp();
    }
}

struct XlWs {
    unsigned char *d1p, *d2p;
    float *n1, *n2;
    unsigned int* ovf;
    float *Kmat, *a, *b, *Tpart;
    int T1, T2, Np8, Mp, upp;
    size_t bytes;
};

XlWs xl_carve(void* ws, int B, int N, int M, int D) {
    XlWs w{};
    w.T1 = (N + XB - 1) / XB;
    w.T2 = (M + XA - 1) / XA;
    w.Np8 = (N + 1 + XS_ROWS - 1) / XS_ROWS * XS_ROWS;
    w.Mp = (M + 1 + 3) / 4 * 4;
    w.upp = (w.Np8 + XS_UNIT * XS_ROWS - 1) / (XS_UNIT * XS_ROWS);
    char* c = (char*)ws;
    w.d1p = (unsigned char*)c; c += align_up((size_t)B * w.T1 * XB * D * 4);
    w.d2p = (unsigned char*)c; c += align_up((size_t)B * w.T2 * XA * D * 4);
    w.n1 = (float*)c; c += align_up((size_t)B * w.T1 * XB * sizeof(float));
    w.n2 = (float*)c; c += align_up((size_t)B * w.T2 * XA * sizeof(float));
    w.ovf = (unsigned int*)c; c += align_up((size_t)B * sizeof(unsigned int));
    w.Kmat = (float*)c; c += align_up((size_t)B * w.Np8 * w.Mp * sizeof(float));
    w.a = (float*)c; c += align_up((size_t)B * w.Np8 * sizeof(float));
    w.b = (float*)c; c += align_up((size_t)B * w.Mp * sizeof(float));
    w.Tpart = (float*)c; c += align_up((size_t)B * w.upp * w.Mp * sizeof(float));
    w.bytes = (size_t)(c - (char*)ws);
    return w;
}

int g_sm_count = 0;
int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = 148;
        }
        g_sm_count = n;
    }
    return g_sm_count;
}

}  // namespace

int g_xl_reverse = 1;               // test / tuning hook: 0 = every sweep walks the units forwards
int g_xl_pdl = 1;                   // test / tuning hook: 0 = plain stream order between the sweep and column kernels
int g_xl_policy = 0;                // tuning hook: L2 policy of the sweeps' matrix reads (l2_policy)

// limits: squared-L2 cost, scaling form in range (as the generic path's test), D % 32 == 0, M + 1 <= XS_MAX_MP
bool sinkhorn_xl_eligible(int N, int M, int D, float eps, float unused, int distance_l1) {
    if (distance_l1 || D <= 0 || D % 32 != 0 || N <= 0 || M <= 0) return false;
    if ((M + 1 + 3) / 4 * 4 > XS_MAX_MP) return false;
    const double dustbin = -(double)unused / (double)eps;
    return dustbin <= 0.0 && -dustbin * 1.4426950408889634 <= 60.0;
}

size_t sinkhorn_xl_workspace_bytes(int B, int N, int M, int D) {
    if (B <= 0 || N <= 0 || M <= 0 || D <= 0) return 0;
    return xl_carve(nullptr, B, N, M, D).bytes;
}

// cost_l2: the FP32 FFMA cost kernel of sinkhorn.cu, writing K with the given pitch for the pairs whose flag is set
int cost_l2_fallback_launch(const float* d1, const float* d2, const float* n1, int n1_stride, const float* n2, int n2_stride, int B,
                            int N, int M, int D, float eps, float dustbin, float* S, int ld, size_t zstride,
                            const unsigned int* per_pair_flag, cudaStream_t st);

// The steps of a call, separately launchable so that the fused matcher can pack each image's descriptors on that image's own
// stream as soon as they exist (as it does for the hybrid kernel): prepare (clears the per-pair range flags), pack (which = 0:
// desc1 rows, 1: desc2 rows), run (cost matrix, iterations, P).
int sinkhorn_xl_prepare(int B, int N, int M, int D, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (ws == nullptr || ws_bytes < sinkhorn_xl_workspace_bytes(B, N, M, D)) return OM_ERR_WORKSPACE;
    const XlWs w = xl_carve(ws, B, N, M, D);
    OM_CUDA(cudaMemsetAsync(w.ovf, 0, (size_t)B * sizeof(unsigned int), st));
    return OM_OK;
}

int sinkhorn_xl_pack(int which, const float* d, int B, int N, int M, int D, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (d == nullptr) return OM_ERR_NULL;
    if (ws == nullptr || ws_bytes < sinkhorn_xl_workspace_bytes(B, N, M, D)) return OM_ERR_WORKSPACE;
    const XlWs w = xl_carve(ws, B, N, M, D);
    if (which == 0) return pack_f16_launch(d, B, N, D, XB, w.T1, XG, w.d1p, w.n1, w.ovf, st);
    return pack_f16_launch(d, B, M, D, XA, w.T2, XG, w.d2p, w.n2, w.ovf, st);
}

int sinkhorn_xl_run(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float eps, float unused,
                    float* P, void* ws, size_t ws_bytes, cudaStream_t st);

int sinkhorn_xl_launch(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float eps, float unused,
                       float* P, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!sinkhorn_xl_eligible(N, M, D, eps, unused, 0)) return OM_ERR_PARAM;
    OM_TRY(sinkhorn_xl_prepare(B, N, M, D, ws, ws_bytes, st));
    OM_TRY(sinkhorn_xl_pack(0, d1, B, N, M, D, ws, ws_bytes, st));
    OM_TRY(sinkhorn_xl_pack(1, d2, B, N, M, D, ws, ws_bytes, st));
    return sinkhorn_xl_run(d1, d2, B, N, M, D, iterations, eps, unused, P, ws, ws_bytes, st);
}

int sinkhorn_xl_run(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float eps, float unused,
                    float* P, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!sinkhorn_xl_eligible(N, M, D, eps, unused, 0)) return OM_ERR_PARAM;
    if (d1 == nullptr || d2 == nullptr || P == nullptr) return OM_ERR_NULL;
    if (ws == nullptr || ws_bytes < sinkhorn_xl_workspace_bytes(B, N, M, D)) return OM_ERR_WORKSPACE;
    const XlWs w = xl_carve(ws, B, N, M, D);
    const size_t zstride = (size_t)w.Np8 * w.Mp;
    const double log2e = 1.4426950408889634;
    const float dustbin = (float)(-(double)unused / (double)eps);          // computed in double by the reference, sinkhorn.py:182
    const int nsm = sm_count();
    {
        CostArgs a{};
        a.d1p = w.d1p; a.d2p = w.d2p; a.n1 = w.n1; a.n2 = w.n2; a.ovf = w.ovf; a.Kmat = w.Kmat;
        a.N = N; a.M = M; a.D = D; a.Mp = w.Mp; a.T1 = w.T1; a.T2 = w.T2; a.total = B * w.T1 * w.T2;
        a.zstride = zstride;
        a.scale2 = (float)(log2e / (double)eps);
        OM_TRY(set_smem(cost_pk_kernel, XC_SMEM));
        cost_pk_kernel<<<a.total < nsm ? a.total : nsm, XC_THREADS, XC_SMEM, st>>>(a);
        OM_AFTER_LAUNCH();
    }
    {
        const int len = w.Mp > N ? w.Mp : N;
        xl_fill_kernel<<<dim3((unsigned)((len + 255) / 256), (unsigned)B), 256, 0, st>>>(w.Kmat, N, M, w.Np8, w.Mp, zstride, expf(dustbin));
        OM_AFTER_LAUNCH();
    }
    // pairs with descriptors beyond the fp16 range (flag set by the packing kernel): FP32 FFMA cost kernel; exits at once otherwise
    OM_TRY(cost_l2_fallback_launch(d1, d2, w.n1, w.T1 * XB, w.n2, w.T2 * XA, B, N, M, D, eps, dustbin, w.Kmat, w.Mp, zstride, w.ovf, st));
    xl_init_b_kernel<<<dim3((unsigned)((w.Mp + 255) / 256), (unsigned)B), 256, 0, st>>>(w.b, M, w.Mp);
    OM_AFTER_LAUNCH();

    SweepArgs s{};
    s.Kmat = w.Kmat; s.b = w.b; s.a = w.a; s.Tpart = w.Tpart;
    s.N = N; s.Np8 = w.Np8; s.Mp = w.Mp; s.upp = w.upp; s.total = B * w.upp; s.zstride = zstride; s.mu_dust = (float)M;
    s.policy = g_xl_policy;
    const size_t ssmem = ((size_t)(XS_SLOTS * (XS_ROWS + 1)) * w.Mp + 16) * 4 + 96;
    OM_TRY(set_smem(xs_sweep_kernel, ssmem));
    const int sgrid = s.total < nsm ? s.total : nsm;
    // sweep and column kernels are launched with programmatic stream serialization: each is scheduled while its
    // predecessor drains and waits (griddepcontrol.wait) before its first dependent access
    cudaLaunchAttribute pdl[1];
    pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl[0].val.programmaticStreamSerializationAllowed = g_xl_pdl ? 1 : 0;
    cudaLaunchConfig_t cs{}, cc{};
    cs.gridDim = dim3((unsigned)sgrid, 1, 1); cs.blockDim = dim3(XS_THREADS, 1, 1); cs.dynamicSmemBytes = ssmem; cs.stream = st;
    cs.attrs = pdl; cs.numAttrs = 1;
    cc.gridDim = dim3((unsigned)((M + 32) / 32), (unsigned)B, 1); cc.blockDim = dim3(256, 1, 1); cc.dynamicSmemBytes = 0; cc.stream = st;
    cc.attrs = pdl; cc.numAttrs = 1;
    for (int it = 0; it < iterations; ++it) {
        s.reverse = g_xl_reverse ? (it & 1) : 0;
        OM_CUDA(cudaLaunchKernelEx(&cs, xs_sweep_kernel, s));
        OM_AFTER_LAUNCH();
        OM_CUDA(cudaLaunchKernelEx(&cc, xs_col_kernel, (const float*)w.Tpart, w.b, w.upp, M, w.Mp, (float)N));
        OM_AFTER_LAUNCH();
    }
    const size_t fsmem = (size_t)8 * w.Mp * sizeof(float);
    OM_TRY(set_smem(xs_finalize_kernel, fsmem));
    int fgrid = (N + 1 + 7) / 8;
    if ((long long)fgrid * B > 8LL * nsm * 3 && fgrid > 1) {           // several rows per warp once the grid covers the GPU a few times
        const int per = (int)((8LL * nsm * 3 + B - 1) / B);
        if (per < fgrid) fgrid = per > 1 ? per : 1;
    }
    xs_finalize_kernel<<<dim3((unsigned)fgrid, (unsigned)B), 256, fsmem, st>>>(w.Kmat, w.a, w.b, P, N, M, w.Np8, w.Mp, zstride);
    OM_AFTER_LAUNCH();
    return OM_OK;
}

}  // namespace om

extern "C" void om_debug_xl_reverse(int on) {
    om::g_xl_reverse = on & 1;
    om::g_xl_pdl = (on & 2) ? 0 : 1;       // bit 1: switch programmatic dependent launch off
    om::g_xl_policy = (on >> 4) & 7;       // bits 4..6: L2 policy of the sweeps
}
