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
#include <cuda_fp16.h>
#include <math_constants.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace om {

namespace {

constexpr int CL = 8;               // CTAs per cluster == per descriptor pair
constexpr int NW = 16;              // producer / sweep warps
constexpr int NP = NW * 32;         // producer threads (stage the GEMM operands, run the Sinkhorn sweep)
constexpr int NT = NP;              // 16 warps: 4 per SM sub-partition -> 128 registers per thread (a 17th warp
                                    // would cap every thread at 96); thread 0 also issues the tcgen05.mma stream
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
constexpr int OFF_CP = OFF_N2 + MAXM;                      // [2][NCOL] this CTA's column partials (bulk-copy source)
constexpr int OFF_BAR = OFF_CP + 2 * NCOL;                 // 8 mbarriers + TMEM base address
constexpr int SMEM_FLOATS = OFF_BAR + 32;
// epilogue tables (fused match extraction): they live in the receive area, which is idle once the iterations are over
constexpr int EPI_COLVAL = OFF_RECV + 1024;                // [CL][64] best value of the owned columns, one row per sender
constexpr int EPI_COLIDX = OFF_RECV + 1536;                // [CL][64] row that attains it
constexpr int EPI_COLARG = OFF_RECV + 2048;                // [MAXM] argmax row of every column (all-gathered)
constexpr int EPI_KEYS = OFF_RECV + 2560;                  // [MAXM] 64-bit sort keys (rank 0)
constexpr int EPI_END = EPI_KEYS + 2 * MAXM;
static_assert(EPI_END <= OFF_CW && (EPI_KEYS * 4) % 8 == 0, "epilogue tables must fit the receive area");
constexpr uint32_t PART_BYTES = NCOL * 4;                  // one rank's partial column sums
static_assert(2 * STAGE_BYTES <= OFF_N1 * 4, "staging must not reach the norms / barriers");
static_assert((OFF_BAR * 4) % 8 == 0 && (OFF_RECV * 4) % 16 == 0 && (OFF_CW * 4) % 16 == 0 && (OFF_CP * 4) % 16 == 0,
              "alignment");

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
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// DSMEM bulk copy: local shared memory -> a peer CTA's shared memory, completion on the PEER's mbarrier
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}
// 4-byte store into a peer CTA's shared memory that also completes 4 bytes on the PEER's mbarrier
__device__ __forceinline__ void st_async_f32(uint32_t dst_cluster, float v, uint32_t bar_cluster) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(dst_cluster),
                 "r"(__float_as_uint(v)), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void st_async_f32x4(uint32_t dst_cluster, const float4& v, uint32_t bar_cluster) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                     dst_cluster), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
                 "r"(__float_as_uint(v.w)), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void bar_sweep() { asm volatile("bar.sync 1, %0;" ::"n"(NP) : "memory"); }
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
// instruction descriptor: D=F32, A=B=F16, both K-major, N=64, M=128
constexpr uint32_t IDESC_F16 = (1u << 4) | (0u << 7) | (0u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc),
        "r"(IDESC_F16), "r"(accumulate) : "memory");
}
// x = hi + lo with hi = fp16(x), lo = fp16(x - hi); four values -> 2 x 8 bytes
__device__ __forceinline__ void split_f16(const float4& v, uint2& hi, uint2& lo) {
    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
    hi.x = *reinterpret_cast<const uint32_t*>(&h01); hi.y = *reinterpret_cast<const uint32_t*>(&h23);
    lo.x = *reinterpret_cast<const uint32_t*>(&l01); lo.y = *reinterpret_cast<const uint32_t*>(&l23);
}
__device__ __forceinline__ float amax4(const float4& v, float m) {
    return fmaxf(fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))), m);
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
    float* P;           // EPI kernels: may be null (P never leaves the chip)
    long long* trace;   // optional (debug): per CTA 8 clock64 stamps
    SinkhornEpilogue e; // EPI kernels only
};

__device__ __forceinline__ void st_async_b64(uint32_t dst_cluster, unsigned long long v, uint32_t bar_cluster) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(dst_cluster), "l"(v),
                 "r"(bar_cluster) : "memory");
}
// float -> unsigned with the same order (so that -1 sorts below every probability); as in matches.cu
__device__ __forceinline__ unsigned int ordered_bits_tc(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits_tc(unsigned int o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

#define OM_STAMP(slot)                                                                   \
    do {                                                                                 \
        if (a.trace != nullptr && threadIdx.x == 0) a.trace[(size_t)blockIdx.x * 8 + (slot)] = clock64(); \
    } while (0)

// XD = true: "scaling" form of the same iteration (see the block comment at the XD branch below)
// F16 = true: operands split into two fp16 terms and multiplied with kind::f16 (K = 16 per MMA) instead of two
// TF32 terms with kind::tf32 (K = 8): same three products hi*hi + hi*lo + lo*hi, a quarter of the MMA instructions
// and half the staged bytes.  x = hi + lo holds 22 bits, so the dot products are as accurate as FP32 FFMA.  fp16
// overflows at 65504: the producers watch max|x| and a CTA that saw |x| >= 60000 recomputes its dots with FFMA.
template <bool XD, bool F16, bool EPI = false>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT, 1) sinkhorn_tc_kernel(TcArgs a) {
    static_assert(!EPI || XD, "the fused epilogue lives in the scaling-form kernel");
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
    float* sCP = sm + OFF_CP;
    // [0,1] stage free (MMAs done), [2] GEMM done, [3,4] partials landed, [5,6] stage full (operands stored)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int r0 = rank * RPC;
    const int nreal = max(0, min(RPC, N - r0));
    const bool has_dust = rank == CL - 1;
    const int nloc = nreal + (has_dust ? 1 : 0);

    // ---------------- one-time setup: mbarriers, TMEM ------------------------------------------
    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        mbar_init(smem_u32(&bars[1]), 1);
        mbar_init(smem_u32(&bars[2]), 1);
        mbar_init(smem_u32(&bars[3]), 1);
        mbar_init(smem_u32(&bars[4]), 1);
        mbar_init(smem_u32(&bars[5]), F16 ? 15 * 32 : NP);        // producers of a stage (F16: warp 15 is the MMA issuer)
        mbar_init(smem_u32(&bars[6]), F16 ? 15 * 32 : NP);
        mbar_init(smem_u32(&bars[7]), 1);                          // epilogue: sort keys landed (rank 0)
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
    OM_STAMP(0);

    // ---------------- similarity GEMM on tcgen05 -------------------------------------------------
    {
        const uint32_t stage0 = smem_u32(sm);
        const float* A = a.d2 + (size_t)z * M * D;                 // MMA "A": 512 rows j of d2
        const float* Bm = a.d1 + ((size_t)z * N + r0) * D;          // MMA "B": this CTA's 64 rows i of d1
        int overflow = 0;                                           // F16: some |x| >= 60000 seen by this CTA
        if constexpr (F16) {
            // ===== fp16 split.  Stage = 32 floats of K (two K=16 MMA steps) =====
            // Tile layout (bytes): element (row, k) at (k/8)*LBO + row*16 + (k%8)*2 -- the same core-matrix
            // geometry as the TF32 tiles (8 rows x 16 bytes, LBO between K groups, SBO = 128 between row groups).
            // Work item = 4 rows x the 32 floats (128 bytes) of the chunk: a warp-level load reads four whole 128-byte
            // lines (lane = row, 16-byte segment) and the two 8-byte stores per lane land in the four K groups of the
            // tile (4 shared-memory wavefronts each).  What paces the producers is L1 wavefronts, one per line touched:
            // the earlier item (16 rows x one K group of 8 floats: 16 half-used lines per load, conflict-free stores)
            // cost 20 wavefronts per 512 bytes staged, this one 12.  144 items per stage (128 of the d2 tile, 16 of the
            // d1 tile) are dealt round-robin to warps 0..14; warp 15 only issues the MMAs, so that staging stage c+1 and
            // the tensor core working on stage c overlap (the tensor pipe paces at ~2.1k cycles per stage).
            constexpr int KC2 = 32, NPW = 15, NITEM = 144, IPW = (NITEM + NPW - 1) / NPW;
            const int nchunks = D / KC2;
            const int prow = lane >> 3, seg = lane & 7;             // row within the item, 16-byte segment of its 128-byte line
            const int kgl = seg >> 1, half = seg & 1;               // K group of 8 floats, half of it
            float mx = 0.f;
            const bool trc = a.trace != nullptr && tid == NPW * 32;
            long long y0 = 0, y1 = 0, y2 = 0, y3 = 0;
            if (warp < NPW) {
                float4 cur[IPW];
                float acc[IPW];
                const float* src[IPW];                              // global address of this thread's float4 in chunk 0 (or null)
                uint32_t dst[IPW];                                  // byte offset of the hi half inside a stage
#pragma unroll
                for (int t = 0; t < IPW; ++t) {
                    const int i = warp + NPW * t;
                    acc[t] = 0.f;
                    src[t] = nullptr;
                    dst[t] = 0;
                    if (i < 128) {                                  // d2 tile (MMA "A")
                        const int j = 4 * i + prow;
                        if (j < M) src[t] = A + (size_t)j * D + 4 * seg;
                        dst[t] = (uint32_t)(kgl * A_LBO + j * 16 + half * 8);
                    } else if (i < NITEM) {                         // d1 tile (MMA "B")
                        const int r = 4 * (i - 128) + prow;
                        if (r < nreal) src[t] = Bm + (size_t)r * D + 4 * seg;
                        dst[t] = (uint32_t)(2 * A_TILE + kgl * B_LBO + r * 16 + half * 8);
                    }
                }
                // items are handled in two halves: as soon as a half is staged its registers are refilled with the next
                // chunk, so those loads are in flight while the other half is being staged (no second register set)
                constexpr int HALF = IPW / 2;
                auto fetch = [&](int c, int t0, int t1) {
#pragma unroll
                    for (int t = 0; t < IPW; ++t)
                        if (t >= t0 && t < t1)
                            cur[t] = (src[t] != nullptr && c < nchunks) ? __ldg(reinterpret_cast<const float4*>(src[t] + c * KC2))
                                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
                };
                auto stage = [&](char* st, int t0, int t1) {
#pragma unroll
                    for (int t = 0; t < IPW; ++t) {
                        const int i = warp + NPW * t;
                        if (t >= t0 && t < t1 && i < NITEM) {       // warp-uniform
                            uint2 hi, lo;
                            acc[t] = sq4(cur[t], acc[t]);
                            mx = amax4(cur[t], mx);
                            split_f16(cur[t], hi, lo);
                            *reinterpret_cast<uint2*>(st + dst[t]) = hi;
                            *reinterpret_cast<uint2*>(st + dst[t] + (i < 128 ? A_TILE : B_TILE)) = lo;
                        }
                    }
                };
                fetch(0, 0, IPW);
                for (int c = 0; c < nchunks; ++c) {
                    const int s = c & 1;
                    if (c >= 2) mbar_wait(smem_u32(&bars[s]), (uint32_t)(((c >> 1) - 1) & 1));   // MMAs of chunk c-2 done
                    char* st = reinterpret_cast<char*>(sm) + s * STAGE_BYTES;
                    stage(st, 0, HALF);
                    fetch(c + 1, 0, HALF);
                    stage(st, HALF, IPW);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async proxy (MMA)
                    mbar_arrive(smem_u32(&bars[5 + s]));                            // stage s is full
                    fetch(c + 1, HALF, IPW);
                }
                // squared norms, sinkhorn.py:98-99: the 8 lanes of a row hold its 8 segments (summed over the chunks)
#pragma unroll
                for (int t = 0; t < IPW; ++t) {
                    const int i = warp + NPW * t;
                    float v = acc[t];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    if (seg == 0 && i < 128) sN2[4 * i + prow] = v;
                    else if (seg == 0 && i < NITEM) sN1[4 * (i - 128) + prow] = v;
                }
            } else if (lane == 0) {
                // ===== MMA issuer (warp 15): wait until the producers have stored chunk c, then drive the tensor core =====
                for (int c = 0; c < nchunks; ++c) {
                    const int s = c & 1;
                    const long long c2 = trc ? clock64() : 0;
                    mbar_wait(smem_u32(&bars[5 + s]), (uint32_t)((c >> 1) & 1));
                    const long long c3 = trc ? clock64() : 0;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = stage0 + s * STAGE_BYTES;
#pragma unroll
                    for (int ks = 0; ks < KC2 / 16; ++ks) {
                        const uint64_t bhi = umma_desc(sa + 2 * A_TILE + 2 * ks * B_LBO, B_LBO, SBO);
                        const uint64_t blo = umma_desc(sa + 2 * A_TILE + B_TILE + 2 * ks * B_LBO, B_LBO, SBO);
#pragma unroll
                        for (int mb = 0; mb < 4; ++mb) {
                            const uint64_t ahi = umma_desc(sa + 2 * ks * A_LBO + mb * 128 * 16, A_LBO, SBO);
                            const uint64_t alo = umma_desc(sa + A_TILE + 2 * ks * A_LBO + mb * 128 * 16, A_LBO, SBO);
                            const uint32_t d = tmem_base + mb * 64;
                            umma_f16(d, ahi, bhi, (c | ks) != 0);
                            umma_f16(d, ahi, blo, 1u);
                            umma_f16(d, alo, bhi, 1u);
                        }
                    }
                    umma_commit(smem_u32(&bars[s]));                    // frees stage s when these MMAs retire
                    if (c == nchunks - 1) umma_commit(smem_u32(&bars[2]));
                    if (trc) { y2 += c3 - c2; y3 += clock64() - c3; }
                }
                if (trc) {
                    long long* y = a.trace + (size_t)gridDim.x * 12 + (size_t)blockIdx.x * 4;
                    y[0] = y0; y[1] = y1; y[2] = y2; y[3] = y3;
                }
            }
            __syncwarp();
            overflow = __syncthreads_or(mx >= 60000.0f ? 1 : 0);    // (also publishes sN1 / sN2)
        } else {
        const int lrow = tid >> 2, lkq = tid & 3;                   // loader: row (+128q), 16-byte K unit
        float nb[4] = {0.f, 0.f, 0.f, 0.f}, na = 0.f;
        const int nchunks = D / KC;
        if (warp < NW) {
            // ===== producers: global -> registers (two chunks ahead) -> hi/lo split -> UMMA layout in smem =====
            float4 pa[4], pb = make_float4(0.f, 0.f, 0.f, 0.f);      // chunk c+1
            float4 qa[4], qb = make_float4(0.f, 0.f, 0.f, 0.f);      // chunk c+2
            auto fetch = [&](int c, float4 (&ra)[4], float4& rb) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = lrow + 128 * q;
                    ra[q] = (j < M && c < nchunks)
                                ? __ldg(reinterpret_cast<const float4*>(A + (size_t)j * D + c * KC + 4 * lkq))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (tid < 4 * B_ROWS)
                    rb = (lrow < nreal && c < nchunks)
                             ? __ldg(reinterpret_cast<const float4*>(Bm + (size_t)lrow * D + c * KC + 4 * lkq))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            fetch(0, pa, pb);
            fetch(1, qa, qb);
            for (int c = 0; c < nchunks; ++c) {
                const int s = c & 1;
                const float4 va[4] = {pa[0], pa[1], pa[2], pa[3]};
                const float4 vb = pb;
#pragma unroll
                for (int q = 0; q < 4; ++q) pa[q] = qa[q];
                pb = qb;
                if (c >= 2) mbar_wait(smem_u32(&bars[s]), (uint32_t)(((c >> 1) - 1) & 1));   // MMAs of chunk c-2 done
                char* st = reinterpret_cast<char*>(sm) + s * STAGE_BYTES;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = lrow + 128 * q;
                    float4 hi, lo;
                    nb[q] = sq4(va[q], nb[q]);
                    split_tf32(va[q], hi, lo);
                    *reinterpret_cast<float4*>(st + lkq * A_LBO + j * 16) = hi;
                    *reinterpret_cast<float4*>(st + A_TILE + lkq * A_LBO + j * 16) = lo;
                }
                if (tid < 4 * B_ROWS) {
                    float4 hi, lo;
                    na = sq4(vb, na);
                    split_tf32(vb, hi, lo);
                    *reinterpret_cast<float4*>(st + 2 * A_TILE + lkq * B_LBO + lrow * 16) = hi;
                    *reinterpret_cast<float4*>(st + 2 * A_TILE + B_TILE + lkq * B_LBO + lrow * 16) = lo;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async proxy (MMA)
                mbar_arrive(smem_u32(&bars[5 + s]));                            // stage s is full
                fetch(c + 2, qa, qb);
                if (tid == 0) {
                    // ===== MMA issue: wait until every producer has stored chunk c, then drive the tensor core =====
                    mbar_wait(smem_u32(&bars[5 + s]), (uint32_t)((c >> 1) & 1));
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
                    umma_commit(smem_u32(&bars[s]));                    // frees stage s when these MMAs retire
                    if (c == nchunks - 1) umma_commit(smem_u32(&bars[2]));
                }

            }
        }
        __syncwarp();
        // squared norms (4 loader threads share a row), sinkhorn.py:98-99
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float v = nb[q];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (lkq == 0 && warp < NW) sN2[lrow + 128 * q] = v;
        }
        na += __shfl_xor_sync(0xffffffffu, na, 1);
        na += __shfl_xor_sync(0xffffffffu, na, 2);
        if (lkq == 0 && tid < 4 * B_ROWS) sN1[lrow] = na;

        }
        OM_STAMP(1);                                                // all chunks staged, last MMAs in flight
        mbar_wait(smem_u32(&bars[2]), 0u);                          // every MMA has completed
        OM_STAMP(2);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        __syncthreads();                                            // norms visible; staging area is dead

        // epilogue: warp (mb, q) reads TMEM lanes 32q..32q+31 of M-block mb: lane = column j, regs = rows i
        const int mb = (warp >> 2) & 3, q = warp & 3;
        const int j = 128 * mb + 32 * q + lane;
        const float n2j = sN2[j];
#pragma unroll
        for (int half = 0; half < 2 && warp < NW; ++half) {
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
        if (F16 && overflow) {                                      // out of fp16 range: plain FP32 dot products (slow, rare)
            for (int li = 0; li < nreal; ++li) {
                const float* x = Bm + (size_t)li * D;
                const float* y = A + (size_t)min(j, M - 1) * D;
                float dot = 0.0f;
                for (int k = 0; k < D; ++k) dot = fmaf(__ldg(x + k), __ldg(y + k), dot);
                const float cost = fmaxf(__fsub_rn(__fadd_rn(sN1[li], n2j), __fmul_rn(2.0f, dot)), 0.0f);
                sS[li * SPITCH + j] = __fmul_rn(-cost, a.scale2);
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
    OM_STAMP(3);                                                    // epilogue done

    if constexpr (XD) {
        // ---------------- Sinkhorn in scaling form, K resident in REGISTERS ----------------------------
        // With K_ij = exp(S_ij), a_i = exp(u_i), b_j = exp(v_j) the log-domain updates of sinkhorn.py:138-142
        //     u_i = log mu_i - LSE_j(S_ij + v_j),   v_j = log nu_j - LSE_i(S_ij + u_i)
        // are  a_i = mu_i / sum_j K_ij b_j,   b_j = nu_j / sum_i K_ij a_i,  and P = a_i K_ij b_j (:145, :206).
        // The launcher takes this path only when exp(-unused/eps) is far from underflow, so every row and
        // column sum keeps a positive dustbin term; K entries may underflow to 0 exactly as exp() of the
        // reference's shifted exponents does.  No exponential is evaluated inside the loop: one sweep is
        // 2 FFMA per matrix entry on registers.  The dustbin row/column are the constant kd and are never stored:
        //     rowsum_i = sum_j K_ij b_j + kd b_M,   rowsum_N = kd (sum_j b_j + b_M)
        //     colsum_j = sum_i K_ij a_i + kd a_N,   colsum_M = kd (sum_i a_i + a_N)
        // Warp w owns score rows 4w..4w+3 of this CTA, lane l columns 4l..4l+3 (+128k): row sums stay inside
        // the warp; column sums go warp -> CTA through shared memory and CTA -> cluster as a reduce-scatter
        // (CTA c owns columns 64c..64c+63, rank 7 also the dustbin column) followed by an all-gather of the
        // new b, both as st.async stores that complete on the receiver's mbarrier (no cluster-wide barrier
        // inside the loop).
        float* sB = sV;                          // b, M+1 entries (entry M = dustbin column)
        float* sPart = sRecv;                    // [CL][72] partial column sums of the owned columns, one row per sender
        float* sAs = sU;                         // per-warp sum of a_i
        const float kd = ex2(a.dustbin2);
        const float Mf = (float)M, Nf = (float)N;
        __syncthreads();                         // score slab complete
        // K as pairs: kreg[rr][2k], kreg[rr][2k+1] = columns 128k+4l+{0,1}, 128k+4l+{2,3} of row 4w+rr
        float2 kreg[4][8];
        if (warp < NW) {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int li = 4 * warp + rr;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c0 = 128 * k + 4 * lane;
                    const float4 s4 = *reinterpret_cast<const float4*>(sS + li * SPITCH + c0);
                    const bool rv = li < nreal;
                    kreg[rr][2 * k].x = (rv && c0 + 0 < M) ? ex2(s4.x) : 0.0f;
                    kreg[rr][2 * k].y = (rv && c0 + 1 < M) ? ex2(s4.y) : 0.0f;
                    kreg[rr][2 * k + 1].x = (rv && c0 + 2 < M) ? ex2(s4.z) : 0.0f;
                    kreg[rr][2 * k + 1].y = (rv && c0 + 3 < M) ? ex2(s4.w) : 0.0f;
                }
            }
        }
        for (int i = tid; i < NCOL; i += NT) sB[i] = i <= M ? 1.0f : 0.0f;      // u = v = 0
        // landing zones in the peers are the same shared-memory offsets, translated per store with mapa:
        // my row of their partial table, their b vector, their two barriers
        const uint32_t my_part = smem_u32(sPart + rank * 72), loc_b = smem_u32(sB);
        const uint32_t loc_bar_part = smem_u32(&bars[3]), loc_bar_b = smem_u32(&bars[4]);
        const uint32_t part_bytes = (uint32_t)CL * (has_dust ? 65u : 64u) * 4u;   // every sender sends all owned slots
        const uint32_t b_bytes = (uint32_t)(CL * 64 + 1) * 4u;
        cluster.sync();      // every CTA is past its GEMM (staging aliased these buffers) and has initialised its barriers
        OM_STAMP(4);

        float2 breg[8];
        float bM = 1.0f, areg[4] = {0.f, 0.f, 0.f, 0.f}, aN = 0.0f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int col = 128 * (c >> 1) + 4 * lane + 2 * (c & 1);
            breg[c] = make_float2(col < M ? 1.0f : 0.0f, col + 1 < M ? 1.0f : 0.0f);
        }
        const bool trx = a.trace != nullptr && tid == 0;
        long long x0 = 0, x1 = 0, x2 = 0, x3 = 0;
        if (warp < NW) {
            for (int it = 0; it < a.iterations; ++it) {
                const uint32_t par = (uint32_t)(it & 1);
                const long long i0 = trx ? clock64() : 0;
                if (tid == 0) {
                    mbar_arrive_expect_tx(smem_u32(&bars[3]), part_bytes);
                    mbar_arrive_expect_tx(smem_u32(&bars[4]), b_bytes);
                }
                // ---- a_i = mu_i / rowsum_i --------------------------------------------------------------
                // (packed fma.rn.f32x2 was measured slower than scalar FFMA here: 1578 vs 1424 cycles per sweep)
                float rs[4];
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    float e0 = 0.0f, e1 = 0.0f;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        e0 = fmaf(kreg[rr][c].x, breg[c].x, e0);
                        e1 = fmaf(kreg[rr][c].y, breg[c].y, e1);
                    }
                    rs[rr] = e0 + e1;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) rs[rr] += __shfl_xor_sync(0xffffffffu, rs[rr], o);
                }
                float asum = 0.0f;
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    areg[rr] = (4 * warp + rr < nreal) ? __fdividef(1.0f, fmaf(kd, bM, rs[rr])) : 0.0f;
                    asum += areg[rr];
                }
                if (warp < 3 || it == a.iterations - 1) {              // only the owner threads (tid <= 64) and the final P need a_N
                    float sb = 0.0f;
#pragma unroll
                    for (int c = 0; c < 8; ++c) sb += breg[c].x + breg[c].y;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) sb += __shfl_xor_sync(0xffffffffu, sb, o);
                    aN = __fdividef(Mf, kd * (sb + bM));               // dustbin row: mu_N = M (sinkhorn.py:197-198)
                }
                // ---- column sums: warp partials -> CTA partials -> owners -------------------------------
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float4 t4;
                    t4.x = fmaf(kreg[3][2 * k].x, areg[3], fmaf(kreg[2][2 * k].x, areg[2], fmaf(kreg[1][2 * k].x, areg[1], kreg[0][2 * k].x * areg[0])));
                    t4.y = fmaf(kreg[3][2 * k].y, areg[3], fmaf(kreg[2][2 * k].y, areg[2], fmaf(kreg[1][2 * k].y, areg[1], kreg[0][2 * k].y * areg[0])));
                    t4.z = fmaf(kreg[3][2 * k + 1].x, areg[3], fmaf(kreg[2][2 * k + 1].x, areg[2], fmaf(kreg[1][2 * k + 1].x, areg[1], kreg[0][2 * k + 1].x * areg[0])));
                    t4.w = fmaf(kreg[3][2 * k + 1].y, areg[3], fmaf(kreg[2][2 * k + 1].y, areg[2], fmaf(kreg[1][2 * k + 1].y, areg[1], kreg[0][2 * k + 1].y * areg[0])));
                    *reinterpret_cast<float4*>(sCW + warp * MAXM + 128 * k + 4 * lane) = t4;
                }
                if (lane == 0) sAs[warp] = asum;
                const long long i1 = trx ? clock64() : 0;
                bar_sweep();
                {   // one column per thread: CTA partial -> the column's owner (measured faster than 128 threads x float4)
                    float s0 = sCW[tid], s1 = sCW[MAXM + tid], s2 = sCW[2 * MAXM + tid], s3 = sCW[3 * MAXM + tid];
#pragma unroll
                    for (int w = 4; w < NW; w += 4) {
                        s0 += sCW[w * MAXM + tid]; s1 += sCW[(w + 1) * MAXM + tid];
                        s2 += sCW[(w + 2) * MAXM + tid]; s3 += sCW[(w + 3) * MAXM + tid];
                    }
                    const uint32_t owner = (uint32_t)tid >> 6;
                    st_async_f32(mapa_u32(my_part + (uint32_t)(tid & 63) * 4u, owner), (s0 + s1) + (s2 + s3),
                                 mapa_u32(loc_bar_part, owner));
                    if (tid == 0) {
                        float as = 0.0f;
#pragma unroll
                        for (int w = 0; w < NW; ++w) as += sAs[w];
                        st_async_f32(mapa_u32(my_part + 64u * 4u, (uint32_t)(CL - 1)), kd * as, mapa_u32(loc_bar_part, (uint32_t)(CL - 1)));
                    }
                }
                // ---- owners: b_j = nu_j / colsum_j, sent to every CTA -----------------------------------
                const long long i2 = trx ? clock64() : 0;
                mbar_wait(smem_u32(&bars[3]), par);
                if (tid < 64 || (tid == 64 && has_dust)) {
                    float t = 0.0f;
#pragma unroll
                    for (int r = 0; r < CL; ++r) t += sPart[r * 72 + tid];
                    const int c = tid == 64 ? M : 64 * rank + tid;      // global column (M = dustbin column)
                    t = fmaf(kd, aN, t);
                    const float bnew = (tid == 64) ? __fdividef(Nf, t) : (c < M ? __fdividef(1.0f, t) : 0.0f);   // nu_M = N (:199-200)
                    const uint32_t slot = (uint32_t)(tid == 64 ? MAXM : c) * 4u;
#pragma unroll
                    for (int dst = 0; dst < CL; ++dst)
                        st_async_f32(mapa_u32(loc_b + slot, (uint32_t)dst), bnew, mapa_u32(loc_bar_b, (uint32_t)dst));
                }
                const long long i3 = trx ? clock64() : 0;
                mbar_wait(smem_u32(&bars[4]), par);
                if (trx) { x0 += i1 - i0; x1 += i2 - i1; x2 += i3 - i2; x3 += clock64() - i3; }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 b4 = *reinterpret_cast<const float4*>(sB + 128 * k + 4 * lane);
                    breg[2 * k] = make_float2(b4.x, b4.y);
                    breg[2 * k + 1] = make_float2(b4.z, b4.w);
                }
                bM = sB[MAXM];
            }
        }
        OM_STAMP(5);
        if (trx) {
            long long* x = a.trace + (size_t)gridDim.x * 8 + (size_t)blockIdx.x * 4;
            x[0] = x0; x[1] = x1; x[2] = x2; x[3] = x3;
        }
        // ---------------- P = a_i K_ij b_j (sinkhorn.py:145, :206) -----------------------------------------
        float* Pz = a.P != nullptr ? a.P + (size_t)z * (N + 1) * (M + 1) : nullptr;
        if constexpr (!EPI) {
        if (warp < NW) {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int li = 4 * warp + rr;
                if (li < nreal) {                                       // warp-uniform
                    float* stage = sS + li * SPITCH;                    // this warp's own slab row: transposes to coalesced stores
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float4 p4;
                        p4.x = areg[rr] * kreg[rr][2 * k].x * breg[2 * k].x;
                        p4.y = areg[rr] * kreg[rr][2 * k].y * breg[2 * k].y;
                        p4.z = areg[rr] * kreg[rr][2 * k + 1].x * breg[2 * k + 1].x;
                        p4.w = areg[rr] * kreg[rr][2 * k + 1].y * breg[2 * k + 1].y;
                        *reinterpret_cast<float4*>(stage + 128 * k + 4 * lane) = p4;
                    }
                    __syncwarp();
                    float* out = Pz + (size_t)(r0 + li) * (M + 1);
#pragma unroll
                    for (int k = 0; k < MAXM / 32; ++k) {
                        const int c = lane + 32 * k;
                        if (c < M) out[c] = stage[c];
                    }
                    if (lane == 0) out[M] = areg[rr] * kd * bM;
                }
            }
            if (has_dust && warp == 0) {                                // dustbin row, b is in shared memory
                float* out = Pz + (size_t)N * (M + 1);
                const float f = aN * kd;
                for (int c = lane; c < M; c += 32) out[c] = f * sB[c];
                if (lane == 0) out[M] = f * bM;
            }
        }
        } else {
        // ======== fused epilogue: filters (sinkhorn.py:311-465), scores (:211-259), mutual matches (match_extraction.py:46-184)
        // on the P values in registers.  P itself is written only when the caller wants it.
        const SinkhornEpilogue& e = a.e;
        const uint32_t epar = (uint32_t)(a.iterations & 1);             // phase of bars[3] / bars[4] after the loop
        float* eWVal = sS;                                              // [NW][MAXM] per-warp column maxima (score slab is dead)
        int* eWIdx = reinterpret_cast<int*>(sS + NW * MAXM);            // [NW][MAXM] local row attaining them
        float* eColVal = sm + EPI_COLVAL;
        int* eColIdx = reinterpret_cast<int*>(sm + EPI_COLIDX);
        int* eColArg = reinterpret_cast<int*>(sm + EPI_COLARG);
        unsigned long long* eKeys = reinterpret_cast<unsigned long long*>(sm + EPI_KEYS);
        const bool want_cols = e.scores1 != nullptr || e.matches;
        if (tid == 0 && want_cols) {
            mbar_arrive_expect_tx(smem_u32(&bars[3]), (uint32_t)CL * 64u * 8u);
            if (e.matches) {
                mbar_arrive_expect_tx(smem_u32(&bars[4]), (uint32_t)MAXM * 4u);
                if (rank == 0) mbar_arrive_expect_tx(smem_u32(&bars[7]), (uint32_t)MAXM * 8u);
            }
        }
        // P in place of K (same expression as the plain kernel's write: bit-identical values)
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                kreg[rr][c].x = areg[rr] * kreg[rr][c].x * breg[c].x;
                kreg[rr][c].y = areg[rr] * kreg[rr][c].y * breg[c].y;
            }
        }
        // ---- row statistics: best / second-best core value and the first column attaining the best ----
        float rbest[4];
        int rarg[4];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            float b1 = -CUDART_INF_F, b2 = -CUDART_INF_F;
            int arg = 0x7fffffff;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int col = 128 * k + 4 * lane + q;
                    const float v = q == 0 ? kreg[rr][2 * k].x : (q == 1 ? kreg[rr][2 * k].y : (q == 2 ? kreg[rr][2 * k + 1].x : kreg[rr][2 * k + 1].y));
                    if (col < M) {
                        if (v > b1) { b2 = b1; b1 = v; arg = col; } else if (v > b2) b2 = v;
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float o1 = __shfl_xor_sync(0xffffffffu, b1, o), o2 = __shfl_xor_sync(0xffffffffu, b2, o);
                const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                b2 = fmaxf(fminf(b1, o1), fmaxf(b2, o2));               // topk(2): a duplicated maximum counts twice
                if (o1 > b1 || (o1 == b1 && oa < arg)) { b1 = o1; arg = oa; }
            }
            const int li = 4 * warp + rr;
            float pdust = areg[rr] * kd * bM;
            bool ok = true;
            if (e.filters) {
                if (e.ratio_threshold > 0.0f) {
                    const float second = M >= 2 ? b2 : 0.0f;            // sinkhorn.py:339-341
                    ok = ok && (__fdiv_rn(b1, __fadd_rn(second, 1e-8f)) >= e.ratio_threshold);
                }
                if (e.dustbin_margin >= 0.0f) ok = ok && (__fsub_rn(b1, pdust) >= e.dustbin_margin);
                if (!ok) {                                              // rejected: core -> 0, dustbin -> 1 (:448-457)
#pragma unroll
                    for (int c = 0; c < 8; ++c) kreg[rr][c] = make_float2(0.0f, 0.0f);
                    pdust = 1.0f;
                    b1 = 0.0f;
                    arg = 0;
                }
                if (lane == 0 && li < nreal && e.filter_valid != nullptr) e.filter_valid[(size_t)z * N + r0 + li] = ok ? 1 : 0;
            }
            rbest[rr] = b1;
            rarg[rr] = arg;
            if (lane == 0 && li < nreal && e.scores0 != nullptr) e.scores0[(size_t)z * N + r0 + li] = b1;
            if (Pz != nullptr && li < nreal) {                          // warp-uniform
                float* stage = sS + li * SPITCH;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    *reinterpret_cast<float4*>(stage + 128 * k + 4 * lane) =
                        make_float4(kreg[rr][2 * k].x, kreg[rr][2 * k].y, kreg[rr][2 * k + 1].x, kreg[rr][2 * k + 1].y);
                __syncwarp();
                float* out = Pz + (size_t)(r0 + li) * (M + 1);
#pragma unroll
                for (int k = 0; k < MAXM / 32; ++k) {
                    const int c = lane + 32 * k;
                    if (c < M) out[c] = stage[c];
                }
                if (lane == 0) out[M] = pdust;
            }
        }
        if (Pz != nullptr && has_dust && warp == 0) {                   // dustbin row: never filtered (:459-460)
            float* out = Pz + (size_t)N * (M + 1);
            const float f = aN * kd;
            for (int c = lane; c < M; c += 32) out[c] = f * sB[c];
            if (lane == 0) out[M] = f * bM;
        }
        if (want_cols) {
            bar_sweep();                                                // every warp is done with its staging rows in sS
            // ---- column statistics: warp partial over its 4 rows -> CTA -> owner of the column ----
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float v4[4];
                int i4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float best = -CUDART_INF_F;
                    int arg = 0;
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {                    // ascending rows, strict >: the first maximal row wins
                        const float v = q == 0 ? kreg[rr][2 * k].x : (q == 1 ? kreg[rr][2 * k].y : (q == 2 ? kreg[rr][2 * k + 1].x : kreg[rr][2 * k + 1].y));
                        if (4 * warp + rr < nreal && v > best) { best = v; arg = 4 * warp + rr; }
                    }
                    v4[q] = best;
                    i4[q] = arg;
                }
                *reinterpret_cast<float4*>(eWVal + warp * MAXM + 128 * k + 4 * lane) = make_float4(v4[0], v4[1], v4[2], v4[3]);
                *reinterpret_cast<int4*>(eWIdx + warp * MAXM + 128 * k + 4 * lane) = make_int4(i4[0], i4[1], i4[2], i4[3]);
            }
            bar_sweep();
            {   // one column per thread: CTA partial -> the column's owner (rows of lower warps first, strict >)
                float best = eWVal[tid];
                int arg = eWIdx[tid];
#pragma unroll
                for (int w = 1; w < NW; ++w) {
                    const float v = eWVal[w * MAXM + tid];
                    if (v > best) { best = v; arg = eWIdx[w * MAXM + tid]; }
                }
                const uint32_t owner = (uint32_t)tid >> 6;
                const uint32_t slot = (uint32_t)(rank * 64 + (tid & 63)) * 4u;
                const uint32_t obar = mapa_u32(smem_u32(&bars[3]), owner);
                st_async_f32(mapa_u32(smem_u32(eColVal) + slot, owner), best, obar);
                st_async_f32(mapa_u32(smem_u32(eColIdx) + slot, owner), __int_as_float(r0 + arg), obar);
            }
            mbar_wait(smem_u32(&bars[3]), epar);
            if (tid < 64) {                                             // owner: best over the 8 row slices, lowest rank first
                float best = eColVal[tid];
                int arg = eColIdx[tid];
#pragma unroll
                for (int r = 1; r < CL; ++r) {
                    const float v = eColVal[r * 64 + tid];
                    if (v > best) { best = v; arg = eColIdx[r * 64 + tid]; }
                }
                const int col = 64 * rank + tid;
                if (col < M && e.scores1 != nullptr) e.scores1[(size_t)z * M + col] = best;
                if (e.matches) {
#pragma unroll
                    for (int dst = 0; dst < CL; ++dst)
                        st_async_f32(mapa_u32(smem_u32(eColArg) + (uint32_t)col * 4u, (uint32_t)dst), __int_as_float(arg),
                                     mapa_u32(smem_u32(&bars[4]), (uint32_t)dst));
                }
            }
        }
        if (e.matches) {
            mbar_wait(smem_u32(&bars[4]), epar);
            // ---- mutual check + threshold (match_extraction.py:95-121): one 64-bit key per row to rank 0 ----
            if (lane < 4) {
                const int rr = lane;
                const int li = 4 * warp + rr;
                const float b1 = rr == 0 ? rbest[0] : (rr == 1 ? rbest[1] : (rr == 2 ? rbest[2] : rbest[3]));
                const int arg = rr == 0 ? rarg[0] : (rr == 1 ? rarg[1] : (rr == 2 ? rarg[2] : rarg[3]));
                unsigned long long key = 0ull;                          // rows beyond N sort last
                if (li < nreal) {
                    const int gi = r0 + li;
                    const bool mutual = eColArg[arg] == gi;
                    const float sc = (mutual && b1 >= e.match_threshold) ? b1 : -1.0f;
                    key = ((unsigned long long)ordered_bits_tc(sc) << 32) | (unsigned long long)(((0x3FFu - (unsigned)gi) << 16) | (unsigned)arg);
                }
                st_async_b64(mapa_u32(smem_u32(eKeys) + (uint32_t)(r0 + li) * 8u, 0u), key, mapa_u32(smem_u32(&bars[7]), 0u));
            }
            if (rank == 0) {
                mbar_wait(smem_u32(&bars[7]), 0u);
                const unsigned long long key0 = eKeys[0];               // row 0's key (its argmax column pads short outputs)
                bar_sweep();
                // descending bitonic sort of the MAXM keys (torch.topk sorted, :124-130; equal scores in ascending row order)
                for (int size = 2; size <= MAXM; size <<= 1) {
                    for (int stride = size >> 1; stride > 0; stride >>= 1) {
                        if (tid < MAXM / 2) {
                            const int lo = 2 * tid - (tid & (stride - 1));
                            const int hi = lo + stride;
                            const bool desc = (lo & size) == 0;
                            const unsigned long long x = eKeys[lo], y = eKeys[hi];
                            if ((x < y) == desc) { eKeys[lo] = y; eKeys[hi] = x; }
                        }
                        bar_sweep();
                    }
                }
                const int take = min(e.max_matches, N);
                for (int m = tid; m < e.max_matches; m += NT) {
                    float sc = 0.0f;                                    // zero padding when N < max_matches (:133-142)
                    unsigned long long key = key0;
                    if (m < take) {
                        key = eKeys[m];
                        sc = from_ordered_bits_tc((unsigned int)(key >> 32));
                    }
                    const int i = m < take ? (int)(0x3FFu - (unsigned)((key >> 16) & 0x3FFu)) : 0;
                    const int j = min((int)(key & 0xFFFFu), M - 1);
                    const size_t o = ((size_t)z * e.max_matches + m) * 2;
                    e.mk1[o] = e.kpts1[((size_t)z * N + i) * 2];
                    e.mk1[o + 1] = e.kpts1[((size_t)z * N + i) * 2 + 1];
                    e.mk2[o] = e.kpts2[((size_t)z * M + j) * 2];
                    e.mk2[o + 1] = e.kpts2[((size_t)z * M + j) * 2 + 1];
                    e.mscores[(size_t)z * e.max_matches + m] = sc;
                    e.mvalid[(size_t)z * e.max_matches + m] = sc > 0.0f ? 1 : 0;      // :181
                }
            }
        }
        }   // EPI
        OM_STAMP(6);
        cluster.sync();      // no CTA may exit while a peer can still write into its shared memory
    } else {
    for (int i = tid; i < 72; i += NT) sU[i] = 0.0f;
    for (int i = tid; i < NCOL; i += NT) sV[i] = 0.0f;
    __syncthreads();

    // ---------------- Sinkhorn iterations (see sinkhorn.cu for the derivation) -----------------------
    const float log2_m = log2f((float)M), log2_n = log2f((float)N);
    float vreg[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) vreg[k] = 0.0f;
    // peers' receive slots for this rank and their "partials landed" barriers, as shared::cluster addresses
    uint32_t peer_slot[CL], peer_bar[CL];
#pragma unroll
    for (int dst = 0; dst < CL; ++dst) {
        peer_slot[dst] = mapa_u32(smem_u32(sRecv + rank * NCOL), (uint32_t)dst);
        peer_bar[dst] = mapa_u32(smem_u32(&bars[3]), (uint32_t)dst);
    }
    // every CTA of the cluster is past its GEMM (its staging aliased sRecv) and has initialised its barriers
    cluster.sync();

    OM_STAMP(4);                                                    // cluster is ready
    long long t_wait = 0, t_sweep = 0, t_red = 0, t_vup = 0;
    const bool tr = a.trace != nullptr;
    for (int it = 0; it < a.iterations; ++it) {
        const int par = it & 1;
        const long long ts0 = tr ? clock64() : 0;
        float colacc[CPL];
#pragma unroll
        for (int k = 0; k < CPL; ++k) colacc[k] = 0.0f;
        for (int li = warp; li < nloc; li += NW) {
            const bool dust_row = has_dust && li == nreal;
            const float mu = dust_row ? (float)M : 1.0f;
            const float* row = sS + li * SPITCH + lane;
            // exp(S_ij + v_j) without the u_i shift: the row factor 2^u_i cancels in mu_i / rowsum, so it is
            // only needed for range -- and the range check below falls back to the max shift when it matters
            float e[CPL];
#pragma unroll
            for (int k = 0; k < CPL; ++k) e[k] = ex2(row[32 * k] + vreg[k]);
            float r0s = e[0], r1s = e[1], r2s = e[2], r3s = e[3];
#pragma unroll
            for (int k = 4; k + 3 < CPL; k += 4) { r0s += e[k]; r1s += e[k + 1]; r2s += e[k + 2]; r3s += e[k + 3]; }
            float rs = ((r0s + r1s) + (r2s + r3s)) + e[CPL - 1];
            rs = warp_sum(rs);
            float u_new, f;
            if (rs >= 1e-30f && rs <= 1e30f) {
                f = __fdividef(mu, rs);
                u_new = (dust_row ? log2_m : 0.0f) - lg2(rs);
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
        const long long ts1 = tr ? clock64() : 0;
        __syncthreads();
        // CTA partial per column -> local staging buffer (double buffered by parity)
        for (int c = tid; c < NCOL; c += NT) {
            float s0 = sCW[c], s1 = sCW[NCOL + c], s2 = sCW[2 * NCOL + c], s3 = sCW[3 * NCOL + c];
#pragma unroll
            for (int w = 4; w < NW; w += 4) {
                s0 += sCW[w * NCOL + c]; s1 += sCW[(w + 1) * NCOL + c];
                s2 += sCW[(w + 2) * NCOL + c]; s3 += sCW[(w + 3) * NCOL + c];
            }
            sCP[par * NCOL + c] = (s0 + s1) + (s2 + s3);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // bulk copy (async proxy) reads it next
        __syncthreads();
        // one thread: arm this CTA's barrier for the 8 incoming partials, then push ours to every CTA of the
        // cluster (incl. ourselves) with DSMEM bulk copies that complete on the receivers' barriers
        if (tid == 0) {
            mbar_arrive_expect_tx(smem_u32(&bars[3 + par]), CL * PART_BYTES);
            const uint32_t src = smem_u32(sCP + par * NCOL);
#pragma unroll
            for (int dst = 0; dst < CL; ++dst)
                bulk_copy_to_peer(peer_slot[dst] + par * CL * PART_BYTES, src, PART_BYTES, peer_bar[dst] + par * 8);
        }
        const long long tw0 = tr ? clock64() : 0;
        mbar_wait(smem_u32(&bars[3 + par]), (uint32_t)((it >> 1) & 1));
        const long long tw1 = tr ? clock64() : 0;
        // v_j += log_nu_j - log(colsum_j)   (sinkhorn.py:142 with the shift -v_j)
        const float* recv = sRecv + par * CL * NCOL;
        for (int c = tid; c <= M; c += NT) {
            float s = 0.0f;
#pragma unroll
            for (int r = 0; r < CL; ++r) s += recv[r * NCOL + c];
            sV[c] = sV[c] + ((c == M ? log2_n : 0.0f) - lg2(s));
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CPL; ++k) vreg[k] = sV[lane + 32 * k];
        if (tr) {
            const long long te = clock64();
            t_sweep += ts1 - ts0; t_red += tw0 - ts1; t_wait += tw1 - tw0; t_vup += te - tw1;
        }
    }

    OM_STAMP(5);                                                    // iterations done
    if (tr && tid == 0) {
        // packed: 4 x 16-bit-shifted averages would lose precision; store wait in slot 7, others after the stamps
        a.trace[(size_t)blockIdx.x * 8 + 7] = t_wait;
        a.trace[(size_t)gridDim.x * 8 + (size_t)blockIdx.x * 4 + 0] = t_sweep;
        a.trace[(size_t)gridDim.x * 8 + (size_t)blockIdx.x * 4 + 1] = t_red;
        a.trace[(size_t)gridDim.x * 8 + (size_t)blockIdx.x * 4 + 2] = t_vup;
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
    OM_STAMP(6);
    cluster.sync();      // no CTA may exit while a peer can still write into its shared memory
    }   // !XD
}


// ------------------------------------------------------------------------------------------
// Cost matrix of the generic path (more than 512 keypoints) on tcgen05: S[z][i][j] for one 64 (i) x 128 (j) tile per CTA,
// same two-term fp16 split, tile layout, producer items and MMA stream as the cluster kernel above (one 128-row block
// of d2 as "A", 64 rows of d1 as "B").  Warps 0..3 stage the operands and read TMEM lanes 32w..32w+31 in the epilogue
// (lane = column j: coalesced rows of S), warp 4 issues the MMAs.  When some |x| >= 60000 was seen (*ovf != 0) the CTA
// exits at once and the FP32 FFMA kernel, launched behind it, does the work instead (and exits at once otherwise).
// ------------------------------------------------------------------------------------------
constexpr int GA_ROWS = 128, GB_ROWS = 64;
constexpr int GA_LBO = GA_ROWS * 16, GB_LBO = GB_ROWS * 16;
constexpr int GA_TERM = 4 * GA_LBO, GB_TERM = 4 * GB_LBO;          // 32 floats of K = 4 K groups of 8 halves
constexpr int G_STAGE = 2 * GA_TERM + 2 * GB_TERM;                 // 24 KB
constexpr int G_THREADS = 160;

__global__ void __launch_bounds__(G_THREADS) cost_tc_kernel(const float* d1, const float* d2, const float* n1, const float* n2,
                                                            int N, int M, int D, float eps, int as_exp, float* S,
                                                            const unsigned int* ovf) {
    extern __shared__ __align__(128) unsigned char gsm[];
    __shared__ __align__(8) uint64_t bars[5];                       // [0,1] stage free, [2] GEMM done, [3,4] stage full
    __shared__ uint32_t tmem_slot;
    if (*ovf != 0u) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int z = blockIdx.z, j0 = blockIdx.x * GA_ROWS, i0 = blockIdx.y * GB_ROWS;
    const float* A = d2 + (size_t)z * M * D;
    const float* Bm = d1 + (size_t)z * N * D;
    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        mbar_init(smem_u32(&bars[1]), 1);
        mbar_init(smem_u32(&bars[2]), 1);
        mbar_init(smem_u32(&bars[3]), 128);
        mbar_init(smem_u32(&bars[4]), 128);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;
    const uint32_t stage0 = smem_u32(gsm);
    constexpr int KC2 = 32, NITEM = 48, IPW = NITEM / 4, HALF = IPW / 2;
    const int nchunks = D / KC2;
    if (warp < 4) {
        const int prow = lane >> 3, seg = lane & 7, kgl = seg >> 1, half = seg & 1;
        float4 cur[IPW];
        const float* src[IPW];
        uint32_t dst[IPW];
#pragma unroll
        for (int t = 0; t < IPW; ++t) {
            const int i = warp + 4 * t;
            src[t] = nullptr;
            if (i < 32) {                                           // d2 rows: MMA "A"
                const int rl = 4 * i + prow, j = j0 + rl;
                if (j < M) src[t] = A + (size_t)j * D + 4 * seg;
                dst[t] = (uint32_t)(kgl * GA_LBO + rl * 16 + half * 8);
            } else {                                                // d1 rows: MMA "B"
                const int rl = 4 * (i - 32) + prow, r = i0 + rl;
                if (r < N) src[t] = Bm + (size_t)r * D + 4 * seg;
                dst[t] = (uint32_t)(2 * GA_TERM + kgl * GB_LBO + rl * 16 + half * 8);
            }
        }
        auto fetch = [&](int c, int t0, int t1) {
#pragma unroll
            for (int t = 0; t < IPW; ++t)
                if (t >= t0 && t < t1)
                    cur[t] = (src[t] != nullptr && c < nchunks) ? __ldg(reinterpret_cast<const float4*>(src[t] + c * KC2))
                                                                : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        auto stage = [&](unsigned char* st, int t0, int t1) {
#pragma unroll
            for (int t = 0; t < IPW; ++t) {
                if (t >= t0 && t < t1) {
                    uint2 hi, lo;
                    split_f16(cur[t], hi, lo);
                    *reinterpret_cast<uint2*>(st + dst[t]) = hi;
                    *reinterpret_cast<uint2*>(st + dst[t] + ((warp + 4 * t) < 32 ? GA_TERM : GB_TERM)) = lo;
                }
            }
        };
        fetch(0, 0, IPW);
        for (int c = 0; c < nchunks; ++c) {
            const int s = c & 1;
            if (c >= 2) mbar_wait(smem_u32(&bars[s]), (uint32_t)(((c >> 1) - 1) & 1));      // MMAs of chunk c-2 done
            unsigned char* st = gsm + s * G_STAGE;
            stage(st, 0, HALF);
            fetch(c + 1, 0, HALF);
            stage(st, HALF, IPW);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(smem_u32(&bars[3 + s]));
            fetch(c + 1, HALF, IPW);
        }
    } else if (lane == 0) {
        for (int c = 0; c < nchunks; ++c) {
            const int s = c & 1;
            mbar_wait(smem_u32(&bars[3 + s]), (uint32_t)((c >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = stage0 + s * G_STAGE;
#pragma unroll
            for (int ks = 0; ks < KC2 / 16; ++ks) {
                const uint64_t ahi = umma_desc(sa + 2 * ks * GA_LBO, GA_LBO, SBO);
                const uint64_t alo = umma_desc(sa + GA_TERM + 2 * ks * GA_LBO, GA_LBO, SBO);
                const uint64_t bhi = umma_desc(sa + 2 * GA_TERM + 2 * ks * GB_LBO, GB_LBO, SBO);
                const uint64_t blo = umma_desc(sa + 2 * GA_TERM + GB_TERM + 2 * ks * GB_LBO, GB_LBO, SBO);
                umma_f16(tmem_base, ahi, bhi, (c | ks) != 0);
                umma_f16(tmem_base, ahi, blo, 1u);
                umma_f16(tmem_base, alo, bhi, 1u);
            }
            umma_commit(smem_u32(&bars[s]));
            if (c == nchunks - 1) umma_commit(smem_u32(&bars[2]));
        }
    }
    __syncwarp();
    mbar_wait(smem_u32(&bars[2]), 0u);                              // every MMA has completed
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
        const int j = j0 + 32 * warp + lane;
        const float n2j = j < M ? n2[(size_t)z * M + j] : 0.0f;
        const float scale2 = (float)(1.4426950408889634 / (double)eps);
        float* Sz = S + (size_t)z * (N + 1) * (M + 1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)(32 * h), r);
#pragma unroll
            for (int ii = 0; ii < 32; ++ii) {
                const int i = i0 + 32 * h + ii;
                if (i < N && j < M) {
                    const float dot = __uint_as_float(r[ii]);
                    const float cost = fmaxf(__fsub_rn(__fadd_rn(n1[(size_t)z * N + i], n2j), __fmul_rn(2.0f, dot)), 0.0f);   // sinkhorn.py:98-103
                    // scaling form: K = exp(-cost / eps) as 2^(-cost * log2(e) / eps), the cluster kernel's arithmetic
                    Sz[(size_t)i * (M + 1) + j] = as_exp ? ex2(__fmul_rn(-cost, scale2)) : __fdiv_rn(-cost, eps);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(64) : "memory");
}

// dustbin row and column of S (sinkhorn.py:182-187) for the tensor-core cost kernel
__global__ void __launch_bounds__(256) dustbin_fill_kernel(float* S, int N, int M, float value) {
    float* Sz = S + (size_t)blockIdx.y * (N + 1) * (M + 1);
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e <= M) Sz[(size_t)N * (M + 1) + e] = value;
    if (e < N) Sz[(size_t)e * (M + 1) + M] = value;
}

}  // namespace

int g_tc_allow_scaling = 1;         // test hook: 0 forces the log-domain loop of the tcgen05 kernel
int g_tc_allow_f16 = 1;             // test hook: 0 forces the 3xTF32 similarity GEMM
long long* g_tc_trace = nullptr;   // debug: device buffer of B*8 CTAs x 8 stamps, set through om_debug_sinkhorn_trace

// tensor-core cost matrix of the generic path; D % 32 == 0.  `ovf` (device): non-zero = out of fp16 range, nothing is done
int cost_tc_launch(const float* d1, const float* d2, const float* n1, const float* n2, int B, int N, int M, int D, float eps,
                   float dustbin, int as_exp, float* S, const unsigned int* ovf, cudaStream_t st) {
    const size_t smem = 2 * (size_t)G_STAGE;     // exactly 48 KB, plus the static barriers: needs the opt-in
    OM_CUDA(cudaFuncSetAttribute(cost_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cost_tc_kernel<<<dim3((M + GA_ROWS - 1) / GA_ROWS, (N + GB_ROWS - 1) / GB_ROWS, B), G_THREADS, smem, st>>>(d1, d2, n1, n2, N, M, D, eps,
                                                                                                               as_exp, S, ovf);
    OM_AFTER_LAUNCH();
    const int len = (N > M ? N : M) + 1;
    dustbin_fill_kernel<<<dim3((len + 255) / 256, B), 256, 0, st>>>(S, N, M, as_exp ? expf(dustbin) : dustbin);
    OM_AFTER_LAUNCH();
    return OM_OK;
}

int sinkhorn_cluster_tc(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float eps,
                        float unused, float* P, cudaStream_t st) {
    TcArgs a{};
    a.trace = g_tc_trace;
    a.d1 = d1; a.d2 = d2; a.N = N; a.M = M; a.D = D; a.iterations = iterations; a.P = P;
    const double log2e = 1.4426950408889634;
    a.scale2 = (float)(log2e / (double)eps);
    a.dustbin2 = (float)((-(double)unused / (double)eps) * log2e);
    const size_t smem = (size_t)SMEM_FLOATS * sizeof(float);
    // scaling form (no exponentials in the loop) while exp(-unused/eps) stays far from underflow (2^-60);
    // beyond that the log-domain loop with its max-shift fallback keeps every sum finite
    const bool scaling = g_tc_allow_scaling && (double)unused / (double)eps * log2e <= 60.0 && unused >= 0.0f;
    const bool f16 = g_tc_allow_f16 && D % 32 == 0;
    if (scaling && f16) {
        OM_TRY(set_smem((sinkhorn_tc_kernel<true, true>), smem));
        sinkhorn_tc_kernel<true, true><<<B * CL, NT, smem, st>>>(a);
    } else if (scaling) {
        OM_TRY(set_smem((sinkhorn_tc_kernel<true, false>), smem));
        sinkhorn_tc_kernel<true, false><<<B * CL, NT, smem, st>>>(a);
    } else if (f16) {
        OM_TRY(set_smem((sinkhorn_tc_kernel<false, true>), smem));
        sinkhorn_tc_kernel<false, true><<<B * CL, NT, smem, st>>>(a);
    } else {
        OM_TRY(set_smem((sinkhorn_tc_kernel<false, false>), smem));
        sinkhorn_tc_kernel<false, false><<<B * CL, NT, smem, st>>>(a);
    }
    OM_AFTER_LAUNCH();
    return OM_OK;
}

bool sinkhorn_epilogue_can_fuse(int N, int M, int D, float eps, float unused, int distance_l1) {
    const double log2e = 1.4426950408889634;
    return !distance_l1 && N <= RPC * CL && M <= MAXM && D % 32 == 0 && g_tc_allow_scaling && g_tc_allow_f16 &&
           (double)unused / (double)eps * log2e <= 60.0 && unused >= 0.0f;
}

int sinkhorn_cluster_tc_epi(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float eps,
                            float unused, float* P, const SinkhornEpilogue& e, cudaStream_t st) {
    if (!sinkhorn_epilogue_can_fuse(N, M, D, eps, unused, 0)) return OM_ERR_PARAM;
    TcArgs a{};
    a.trace = nullptr;
    a.d1 = d1; a.d2 = d2; a.N = N; a.M = M; a.D = D; a.iterations = iterations; a.P = P; a.e = e;
    const double log2e = 1.4426950408889634;
    a.scale2 = (float)(log2e / (double)eps);
    a.dustbin2 = (float)((-(double)unused / (double)eps) * log2e);
    const size_t smem = (size_t)SMEM_FLOATS * sizeof(float);
    OM_TRY(set_smem((sinkhorn_tc_kernel<true, true, true>), smem));
    sinkhorn_tc_kernel<true, true, true><<<B * CL, NT, smem, st>>>(a);
    OM_AFTER_LAUNCH();
    return OM_OK;
}

}  // namespace om

extern "C" void om_debug_sinkhorn_trace(long long* device_buffer) { om::g_tc_trace = device_buffer; }
