// Matching stage, Blackwell-native variant: one 8-CTA cluster per descriptor pair, similarity GEMM on
// the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM), score matrix
// resident in distributed shared memory for all Sinkhorn iterations.
//
// Replaces matching/sinkhorn.py:149-208 (same arithmetic contract as sinkhorn.cu, see there for the
// log-domain formulation and why one exp per entry per iteration suffices).
//
// GEMM.  CTA `rank` owns score rows i in [64*rank, 64*rank+64) and needs G[i][j] = d1_i . d2_j for
// all j < 512.  It is computed transposed, D[j][i] = sum_k d2[j][k] * d1[i][k], so that the MMA runs
// with M = 128 (full tensor-core rate; M = 64 would halve it): 4 M-blocks of 128 columns j, N = 64
// rows i, K = 8 per instruction.  TF32 alone misses the 1e-4 tolerance at epsilon = 0.05 (SURVEY.md
// E10), so every operand is split x = hi + lo with hi the TF32 rounding of x, and three products
// hi*hi + hi*lo + lo*hi are accumulated in the same FP32 TMEM accumulator (error ~2^-21, FP32-like).
// Both operands are K-major in global memory already (descriptor rows); threads stage them into the
// no-swizzle canonical UMMA layout (8-row x 16-byte core matrices) with plain 16-byte stores, two
// stages deep, and one thread issues the MMAs; tcgen05.commit -> mbarrier frees a stage.  The staging
// buffers alias the shared memory that holds the score slab afterwards.
// Epilogue: tcgen05.ld (one TMEM lane = one column j per thread) -> cost -> base-2 log score -> smem.
#include <cooperative_groups.h>
#include <math_constants.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace om {

namespace {

constexpr int CL = 8;               // CTAs per cluster == per descriptor pair
constexpr int NT = 512;             // threads per CTA
constexpr int NW = NT / 32;
constexpr int RPC = 64;             // real score rows per CTA
constexpr int MAXM = 512;
constexpr int NCOL = 544;           // 17 columns per lane; columns beyond M hold -inf
constexpr int CPL = NCOL / 32;
constexpr int SPITCH = NCOL;
constexpr int KC = 16;              // floats of K staged per pipeline stage (two K=8 MMA steps)

// staging geometry, bytes.  Element (row, k) of a tile lives at (k/4)*LBO + row*16 + (k%4)*4:
// 8 consecutive rows form one 128-byte core matrix (SBO = 128), 16-byte K units are LBO apart.
constexpr int A_ROWS = 512, B_ROWS = 64;
constexpr int A_LBO = A_ROWS * 16, B_LBO = B_ROWS * 16, SBO = 128;
constexpr int A_TILE = (KC / 4) * A_LBO;     // 32 KB: one of {hi, lo} of the d2 chunk
constexpr int B_TILE = (KC / 4) * B_LBO;     //  4 KB: one of {hi, lo} of the d1 chunk
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;
constexpr int TMEM_COLS = 256;               // 4 M-blocks x 64 fp32 columns

// shared-memory plan, in floats.  [0, 2*STAGE_BYTES/4) doubles as the GEMM staging area.
constexpr int OFF_S = 0;                                   // (RPC+1) x SPITCH score slab
constexpr int OFF_U = OFF_S + (RPC + 1) * SPITCH;          // u (base 2) per local row
constexpr int OFF_V = OFF_U + 72;                          // v (base 2) per column
constexpr int OFF_RECV = OFF_V + NCOL;                     // [2][CL][NCOL] column partials from the cluster
constexpr int OFF_CW = OFF_RECV + 2 * CL * NCOL;           // [NW][NCOL] per-warp column accumulators
constexpr int OFF_N1 = OFF_CW + NW * NCOL;                 // squared norms of the local d1 rows
constexpr int OFF_N2 = OFF_N1 + 64;                        // squared norms of all d2 rows
constexpr int OFF_BAR = OFF_N2 + MAXM;                     // 3 mbarriers + TMEM base address
constexpr int SMEM_FLOATS = OFF_BAR + 8;
static_assert(2 * STAGE_BYTES <= OFF_N1 * 4, "staging must not reach the norms / barriers");
static_assert((OFF_BAR * 4) % 8 == 0 && (OFF_RECV * 4) % 16 == 0 && (OFF_CW * 4) % 16 == 0, "alignment");

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// UMMA shared-memory descriptor, no swizzle, K-major (cute/arch/mma_sm100_desc.hpp layout)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D=F32, A=B=TF32, both K-major, N=64, M=128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc),
        "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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

// x = hi + lo with hi = x rounded to TF32 (10 explicit mantissa bits); lo is exact in fp32
__device__ __forceinline__ void split_tf32(const float4& v, float4& hi, float4& lo) {
    hi.x = __uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xFFFFE000u);
    hi.y = __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xFFFFE000u);
    hi.z = __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xFFFFE000u);
    hi.w = __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xFFFFE000u);
    lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
}
__device__ __forceinline__ float sq4(const float4& v, float acc) {
    return fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, acc))));
}

struct TcArgs {
    const float* d1;
    const float* d2;
    int N, M, D;
    int iterations;
    float scale2;       // log2(e)/eps
    float dustbin2;     // (-unused/eps) * log2(e)
    float* P;
};

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT, 1) sinkhorn_tc_kernel(TcArgs a) {
    extern __shared__ __align__(128) float sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int z = blockIdx.x / CL;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = a.N, M = a.M, D = a.D;

    float* sS = sm + OFF_S;
    float* sU = sm + OFF_U;
    float* sV = sm + OFF_V;
    float* sRecv = sm + OFF_RECV;
    float* sCW = sm + OFF_CW;
    float* sN1 = sm + OFF_N1;
    float* sN2 = sm + OFF_N2;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

    const int r0 = rank * RPC;
    const int nreal = max(0, min(RPC, N - r0));
    const bool has_dust = rank == CL - 1;
    const int nloc = nreal + (has_dust ? 1 : 0);

    // ---------------- one-time setup: mbarriers, TMEM ------------------------------------------
    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        mbar_init(smem_u32(&bars[1]), 1);
        mbar_init(smem_u32(&bars[2]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // ---------------- similarity GEMM on tcgen05 -------------------------------------------------
    {
        const uint32_t stage0 = smem_u32(sm);
        const float* A = a.d2 + (size_t)z * M * D;                 // MMA "A": 512 rows j of d2
        const float* Bm = a.d1 + ((size_t)z * N + r0) * D;          // MMA "B": this CTA's 64 rows i of d1
        const int lrow = tid >> 2, lkq = tid & 3;                   // loader: row (+128q), 16-byte K unit
        float nb[4] = {0.f, 0.f, 0.f, 0.f}, na = 0.f;
        const int nchunks = D / KC;
        for (int c = 0; c < nchunks; ++c) {
            const int s = c & 1;
            if (c >= 2) mbar_wait(smem_u32(&bars[s]), (uint32_t)(((c >> 1) - 1) & 1));   // MMAs of chunk c-2 done
            char* st = reinterpret_cast<char*>(sm) + s * STAGE_BYTES;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = lrow + 128 * q;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f), hi, lo;
                if (j < M) v = __ldg(reinterpret_cast<const float4*>(A + (size_t)j * D + c * KC + 4 * lkq));
                nb[q] = sq4(v, nb[q]);
                split_tf32(v, hi, lo);
                *reinterpret_cast<float4*>(st + lkq * A_LBO + j * 16) = hi;
                *reinterpret_cast<float4*>(st + A_TILE + lkq * A_LBO + j * 16) = lo;
            }
            if (tid < 4 * B_ROWS) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f), hi, lo;
                if (lrow < nreal) v = __ldg(reinterpret_cast<const float4*>(Bm + (size_t)lrow * D + c * KC + 4 * lkq));
                na = sq4(v, na);
                split_tf32(v, hi, lo);
                *reinterpret_cast<float4*>(st + 2 * A_TILE + lkq * B_LBO + lrow * 16) = hi;
                *reinterpret_cast<float4*>(st + 2 * A_TILE + B_TILE + lkq * B_LBO + lrow * 16) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async proxy (MMA)
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = stage0 + s * STAGE_BYTES;
#pragma unroll
                for (int ks = 0; ks < KC / 8; ++ks) {
                    const uint64_t bhi = umma_desc(sa + 2 * A_TILE + 2 * ks * B_LBO, B_LBO, SBO);
                    const uint64_t blo = umma_desc(sa + 2 * A_TILE + B_TILE + 2 * ks * B_LBO, B_LBO, SBO);
#pragma unroll
                    for (int mb = 0; mb < 4; ++mb) {
                        const uint64_t ahi = umma_desc(sa + 2 * ks * A_LBO + mb * 128 * 16, A_LBO, SBO);
                        const uint64_t alo = umma_desc(sa + A_TILE + 2 * ks * A_LBO + mb * 128 * 16, A_LBO, SBO);
                        const uint32_t d = tmem_base + mb * 64;
                        umma_tf32(d, ahi, bhi, (c | ks) != 0);
                        umma_tf32(d, ahi, blo, 1u);
                        umma_tf32(d, alo, bhi, 1u);
                    }
                }
                umma_commit(smem_u32(&bars[s]));
                if (c == nchunks - 1) umma_commit(smem_u32(&bars[2]));
            }
        }
        // squared norms (4 loader threads share a row), sinkhorn.py:98-99
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float v = nb[q];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (lkq == 0) sN2[lrow + 128 * q] = v;
        }
        na += __shfl_xor_sync(0xffffffffu, na, 1);
        na += __shfl_xor_sync(0xffffffffu, na, 2);
        if (lkq == 0 && tid < 4 * B_ROWS) sN1[lrow] = na;

        mbar_wait(smem_u32(&bars[2]), 0u);                          // every MMA has completed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        __syncthreads();                                            // norms visible; staging area is dead

        // epilogue: warp (mb, q) reads TMEM lanes 32q..32q+31 of M-block mb: lane = column j, regs = rows i
        const int mb = warp >> 2, q = warp & 3;
        const int j = 128 * mb + 32 * q + lane;
        const float n2j = sN2[j];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(mb * 64 + 32 * half), r);
#pragma unroll
            for (int ii = 0; ii < 32; ++ii) {
                const int li = 32 * half + ii;
                if (li < nreal) {
                    const float dot = __uint_as_float(r[ii]);
                    const float cost = fmaxf(__fsub_rn(__fadd_rn(sN1[li], n2j), __fmul_rn(2.0f, dot)), 0.0f);   // :98-103
                    sS[li * SPITCH + j] = __fmul_rn(-cost, a.scale2);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        // dustbin column / row (sinkhorn.py:182-187); columns beyond M are -inf so they drop out of every sum
        for (int e2 = tid; e2 < nloc * (NCOL - M); e2 += NT) {
            const int li = e2 / (NCOL - M), jj = M + e2 % (NCOL - M);
            sS[li * SPITCH + jj] = (jj == M) ? a.dustbin2 : -CUDART_INF_F;
        }
        if (has_dust)
            for (int jj = tid; jj < M; jj += NT) sS[nreal * SPITCH + jj] = a.dustbin2;
    }
    for (int i = tid; i < 72; i += NT) sU[i] = 0.0f;
    for (int i = tid; i < NCOL; i += NT) sV[i] = 0.0f;
    __syncthreads();

    // ---------------- Sinkhorn iterations (see sinkhorn.cu for the derivation) -----------------------
    const float log2_m = log2f((float)M), log2_n = log2f((float)N);
    float vreg[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) vreg[k] = 0.0f;
    float* peer_slot[CL];
#pragma unroll
    for (int dst = 0; dst < CL; ++dst) peer_slot[dst] = cluster.map_shared_rank(sRecv + rank * NCOL, dst);
    cluster.sync();      // every CTA of the cluster is past its GEMM (its staging aliased sRecv) before any DSMEM store

    for (int it = 0; it < a.iterations; ++it) {
        float colacc[CPL];
#pragma unroll
        for (int k = 0; k < CPL; ++k) colacc[k] = 0.0f;
        for (int li = warp; li < nloc; li += NW) {
            const bool dust_row = has_dust && li == nreal;
            const float mu = dust_row ? (float)M : 1.0f;
            const float u_old = sU[li];
            const float* row = sS + li * SPITCH + lane;
            float e[CPL];
            float rs = 0.0f;
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                e[k] = ex2((row[32 * k] + vreg[k]) + u_old);
                rs += e[k];
            }
            rs = warp_sum(rs);
            float u_new, f;
            if (rs >= 1e-30f && rs <= 1e30f) {
                f = __fdividef(mu, rs);
                u_new = u_old + ((dust_row ? log2_m : 0.0f) - lg2(rs));
            } else {
                float m = -CUDART_INF_F;                              // classic max shift, sinkhorn.py:140
#pragma unroll
                for (int k = 0; k < CPL; ++k) m = fmaxf(m, row[32 * k] + vreg[k]);
                m = warp_max(m);
                rs = 0.0f;
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    e[k] = ex2((row[32 * k] + vreg[k]) - m);
                    rs += e[k];
                }
                rs = warp_sum(rs);
                f = __fdividef(mu, rs);
                u_new = (dust_row ? log2_m : 0.0f) - (m + lg2(rs));
            }
#pragma unroll
            for (int k = 0; k < CPL; ++k) colacc[k] = fmaf(e[k], f, colacc[k]);
            if (lane == 0) sU[li] = u_new;
        }
#pragma unroll
        for (int k = 0; k < CPL; ++k) sCW[warp * NCOL + lane + 32 * k] = colacc[k];
        __syncthreads();
        if (tid < NCOL / 4) {
            float4 s4 = *reinterpret_cast<const float4*>(sCW + 4 * tid);
#pragma unroll
            for (int w = 1; w < NW; ++w) {
                const float4 t4 = *reinterpret_cast<const float4*>(sCW + w * NCOL + 4 * tid);
                s4.x += t4.x; s4.y += t4.y; s4.z += t4.z; s4.w += t4.w;
            }
            const int off = (it & 1) * CL * NCOL + 4 * tid;
#pragma unroll
            for (int dst = 0; dst < CL; ++dst) *reinterpret_cast<float4*>(peer_slot[dst] + off) = s4;
        }
        cluster.sync();
        const float* recv = sRecv + (it & 1) * CL * NCOL;
        for (int c = tid; c <= M; c += NT) {
            float s = 0.0f;
#pragma unroll
            for (int r = 0; r < CL; ++r) s += recv[r * NCOL + c];
            sV[c] = sV[c] + ((c == M ? log2_n : 0.0f) - lg2(s));
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CPL; ++k) vreg[k] = sV[lane + 32 * k];
    }

    // P = exp(S + u + v) (sinkhorn.py:145, :206)
    float* Pz = a.P + (size_t)z * (N + 1) * (M + 1);
    for (int li = warp; li < nloc; li += NW) {
        const bool dust_row = has_dust && li == nreal;
        const int gi = dust_row ? N : r0 + li;
        const float u = sU[li];
        const float* row = sS + li * SPITCH;
        float* out = Pz + (size_t)gi * (M + 1);
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const int c = lane + 32 * k;
            if (c <= M) out[c] = ex2((row[c] + u) + vreg[k]);
        }
    }
    cluster.sync();      // no CTA may exit while a peer can still write into its shared memory
}

}  // namespace

int sinkhorn_cluster_tc(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float eps,
                        float unused, float* P, cudaStream_t st) {
    TcArgs a{};
    a.d1 = d1; a.d2 = d2; a.N = N; a.M = M; a.D = D; a.iterations = iterations; a.P = P;
    const double log2e = 1.4426950408889634;
    a.scale2 = (float)(log2e / (double)eps);
    a.dustbin2 = (float)((-(double)unused / (double)eps) * log2e);
    const size_t smem = (size_t)SMEM_FLOATS * sizeof(float);
    OM_TRY(set_smem(sinkhorn_tc_kernel, smem));
    sinkhorn_tc_kernel<<<B * CL, NT, smem, st>>>(a);
    OM_AFTER_LAUNCH();
    return OM_OK;
}

}  // namespace om
