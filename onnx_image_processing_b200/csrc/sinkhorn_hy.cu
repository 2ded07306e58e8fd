// Matching stage, "hybrid-resident" cluster kernel: the whole (N x M) kernel matrix K = exp(-cost/eps) of a descriptor
// pair stays on the chip for all Sinkhorn iterations, part of every CTA's rows in REGISTERS and the rest in TENSOR MEMORY
// (used as a per-thread scratchpad once the similarity GEMM's accumulators are dead: 128 private words per thread, read
// at ~800 B/clk per SM against 128 B/clk for shared memory), so that one SM holds twice the rows of sinkhorn_tc_kernel
// (sinkhorn_tc.cu) and a pair needs half the SMs:
//
//   * up to 512 x 512:   4-CTA cluster, 128 rows per CTA (per thread 48 values in registers + 80 in tensor memory)
//     -> 33 pairs in flight on a B200 instead of ~14 with the 8-CTA kernel, and half the cross-CTA exchange partners;
//   * up to 1024 x 1024 (the reference's export defaults, onnx_export/export_shi_tomasi_sparse_bad_sinkhorn.py:52-127):
//     16-CTA (non-portable) cluster, 64 rows x 1024 columns per CTA (64 + 64 values per thread; 7 clusters resident) --
//     the matrix no longer makes 20 round trips through global memory as on the generic path.
// History of the K <= 512 form, 64 pairs: K rows in shared memory, two-hop exchange 250 us -> no spilled K values, one-hop
// exchange 193 us -> K rows in tensor memory 169 us (K = 1024: 1405 -> 1047 us).
//
// Replaces matching/sinkhorn.py:79-208 in the scaling form (a = mu / (K b), b = nu / (K^T a), P = a K b; see the block
// comment in sinkhorn_tc.cu for the derivation and its range condition).
//
// Similarity GEMM: tcgen05.mma kind::f16 on operands split into two fp16 terms (hi*hi + hi*lo + lo*hi, FP32 accumulators
// in all 512 TMEM columns).  The terms are produced ONCE per descriptor row by pack_f16_kernel, already in the
// no-swizzle K-major core-matrix layout the MMA reads, so a stage of the pipeline is two 1-D bulk copies (UBLKCP)
// issued by one thread; there are no producer warps and no per-CTA re-conversion of d2.
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <math_constants.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace om {

namespace {

// 4-CTA form: rows per warp (of 8) whose K values live in registers; the others live in tensor memory (3 and 4 measure the
// same; with the other rows in SHARED memory 4 rows spilled 20 values per thread: 244 us against 202 us for 3 rows)
#define HY_RR4 3
constexpr int NW = 16;              // warps per CTA
constexpr int NT = NW * 32;

template <int CL_, int MAXM_, int RR_>
struct Hy {
    static constexpr int CL = CL_;                   // CTAs per cluster == per descriptor pair
    static constexpr int MAXM = MAXM_;               // columns (keypoints of image 2) at most
    static constexpr int NK = MAXM / 128;            // float4 column groups per lane: lane l owns columns 128k + 4l .. +3
    static constexpr int RPW = 32 / NK;              // rows per warp (128 K values per thread)
    static constexpr int RR = RR_;                   // rows per warp kept in registers (16 RR / ... K values per thread)
    static constexpr int RS = RPW - RR;              // rows per warp kept in shared memory
    static constexpr int RPC = NW * RPW;             // rows per CTA: 128 (MAXM 512) / 64 (MAXM 1024)
    static constexpr int OWN = MAXM / CL;            // columns whose sums this CTA finishes: 128 / 64
    static constexpr int PSTR = OWN + 8;             // pitch of the partial table (slot OWN: dustbin column, last rank)
    static constexpr int KC = MAXM == 512 ? 32 : 16; // descriptor elements per pipeline stage
    static constexpr int G = KC / 8;                 // 16-byte K groups per stage
    static constexpr int A_TERM = G * MAXM * 16;     // bytes of one fp16 term of a full d2 stage (32 KB)
    static constexpr int B_TERM = G * RPC * 16;      // bytes of one fp16 term of this CTA's d1 stage (8 KB / 2 KB)
    static constexpr int STAGE = 2 * A_TERM + 2 * B_TERM;
    static constexpr int NMB = MAXM / 128;           // M blocks of the (transposed) GEMM
    static constexpr int TMEM_COLS = 512;            // NMB * RPC
    // shared memory, in floats
    static constexpr int OFF_KS = 0;                               // [NW][RS][MAXM] rows of K kept in shared memory
    static constexpr int OFF_CW = OFF_KS + NW * RS * MAXM;         // [NW][MAXM] per-warp column partials | bounce | P staging
    static constexpr int OFF_B = OFF_CW + NW * MAXM;               // b: MAXM entries + dustbin column at [MAXM]
    static constexpr int OFF_PART = OFF_B + MAXM + 32;             // [CL][PSTR] partial column sums of the owned columns
    static constexpr int OFF_AS = OFF_PART + CL * PSTR;            // per-warp sum of a
    static constexpr int OFF_N1 = OFF_AS + 32;                     // squared norms of the local d1 rows
    static constexpr int OFF_S1 = OFF_N1 + RPC;                    // binary operands: the value of a row's non-zero entries (else 1)
    static constexpr int OFF_MISC = OFF_S1 + RPC;                  // fused epilogue, written by peers: column argmax [MAXM],
                                                                   // (value, row) words of the owned columns [CL][OWN] x 8 B,
                                                                   // sort keys [CL * RPC] x 8 B (rank 0)
    // all-to-all exchange of the column partials (small clusters only): [2 buffers][CL ranks][XSTR]; slot MAXM of a rank's
    // row carries its share of the dustbin column
    static constexpr bool A2A = CL <= 4;
    static constexpr int XSTR = MAXM + 8;
    static constexpr int OFF_X = OFF_MISC + 5 * MAXM + 64;
    static constexpr int OFF_BAR = OFF_X + (A2A ? 2 * CL * XSTR : 0);   // 10 mbarriers + TMEM base address
    static constexpr int SMEM_FLOATS = OFF_BAR + 32;
    static_assert(NMB * RPC == TMEM_COLS, "accumulators fill the tensor memory");
    static_assert(2 * STAGE <= OFF_B * 4, "GEMM staging must alias only the K rows and the column partials");
    static_assert((OFF_BAR * 4) % 8 == 0 && (OFF_MISC * 4) % 8 == 0 && (OFF_B * 4) % 16 == 0 && (OFF_CW * 4) % 16 == 0 &&
                  (OFF_X * 4) % 16 == 0 && (XSTR * 4) % 16 == 0, "alignment");
    static_assert(CL * PSTR >= MAXM, "the partial table doubles as the epilogue's column-row table");
    static_assert(SMEM_FLOATS * 4 <= 227 * 1024, "shared memory");
};

__device__ __forceinline__ float ex2h(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t mapa32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async32(uint32_t dst_cluster, float v, uint32_t bar_cluster) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(dst_cluster),
                 "r"(__float_as_uint(v)), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void st_async64(uint32_t dst_cluster, unsigned long long v, uint32_t bar_cluster) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(dst_cluster), "l"(v),
                 "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t dst_cluster, const float4& v, uint32_t bar_cluster) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst_cluster),
                 "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)),
                 "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void bar_all() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }
// 1-D bulk copy global -> this CTA's shared memory, completion (bytes) on a local mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar) : "memory");
}
// the same to the same shared-memory offset of every CTA in `mask` (completion on each one's mbarrier at that offset)
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                     dst_smem), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
// UMMA shared-memory descriptor, no swizzle, K-major (cute/arch/mma_sm100_desc.hpp layout)
__device__ __forceinline__ uint64_t udesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}
__device__ __forceinline__ void umma_f16_n(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(accumulate) : "memory");
}
// the same on 8-bit operands (E4M3: 0x38 = 1.0, 0x00 = 0.0), FP32 accumulators: K = 32 elements per instruction
__device__ __forceinline__ void umma_f8_n(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_to(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// arrives on the mbarrier at this offset in every CTA of `mask` once the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld32x32(uint32_t taddr, uint32_t (&r)[32]) {
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
// tensor memory as a per-thread scratchpad: four 32-bit columns of the calling thread's own lane (lane quadrant of its warp)
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float4& v) {
    uint32_t a, b, c, d;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr));
    v.x = __uint_as_float(a); v.y = __uint_as_float(b); v.z = __uint_as_float(c); v.w = __uint_as_float(d);
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float4& v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(v.x)),
                 "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// packed f32x2 forms (FFMA2): two lanes of FP32 work per issue slot -- the FP32 rate is the same, the issue slots halve
__device__ __forceinline__ float2 dot4x2(const float4& k, const float4& b, float2 acc) {
    acc = __ffma2_rn(make_float2(k.x, k.y), make_float2(b.x, b.y), acc);
    return __ffma2_rn(make_float2(k.z, k.w), make_float2(b.z, b.w), acc);
}
__device__ __forceinline__ void axpy4x2(const float4& k, float2 a2, float2& lo, float2& hi) {
    lo = __ffma2_rn(make_float2(k.x, k.y), a2, lo);
    hi = __ffma2_rn(make_float2(k.z, k.w), a2, hi);
}
__device__ __forceinline__ float dot4(const float4& k, const float4& b, float acc) {
    return fmaf(k.w, b.w, fmaf(k.z, b.z, fmaf(k.y, b.y, fmaf(k.x, b.x, acc))));
}
__device__ __forceinline__ unsigned int ordered_bits_hy(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits_hy(unsigned int o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// ------------------------------------------------------------------------------------------
// Operand packing: descriptor rows (float32) -> two fp16 terms in the MMA's core-matrix layout, squared norms, range flag.
//   out[z][tile][chunk][term][g][Rp rows][8 halves]      tile = row / Rp (d1: one tile per CTA of the cluster; d2: one tile)
// A warp converts 8 rows x 32 elements per step: it reads eight whole 128-byte lines and writes 128-byte core matrices.
// Rows beyond `rows` (padding up to tiles * Rp) are written as zeros.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_f16_kernel(const float* d, int rows, int D, int Rp, int tiles, int G,
                                                       unsigned char* out, float* norms, unsigned int* ovf) {
    const int z = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * 8 + (threadIdx.x >> 5);          // 8 consecutive padded rows
    const int prow = item * 8 + (lane >> 2);
    if (item * 8 >= tiles * Rp) return;
    const int g4 = lane & 3;
    const int tile = prow / Rp, rl = prow - tile * Rp;
    const int nchunks = D / (8 * G);
    const bool real = prow < rows;
    const float* src = d + ((size_t)z * rows + (real ? prow : 0)) * D + 8 * g4;
    uint4* dst = reinterpret_cast<uint4*>(out) + (size_t)z * tiles * nchunks * 2 * G * Rp;
    float acc = 0.0f, mx = 0.0f;
    for (int c32 = 0; c32 < D / 32; ++c32) {
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (real) {
            v0 = __ldg(reinterpret_cast<const float4*>(src + 32 * c32));
            v1 = __ldg(reinterpret_cast<const float4*>(src + 32 * c32 + 4));
        }
        acc = fmaf(v0.x, v0.x, fmaf(v0.y, v0.y, fmaf(v0.z, v0.z, fmaf(v0.w, v0.w, acc))));
        acc = fmaf(v1.x, v1.x, fmaf(v1.y, v1.y, fmaf(v1.z, v1.z, fmaf(v1.w, v1.w, acc))));
        mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(fabsf(v0.x), fabsf(v0.y)), fmaxf(fabsf(v0.z), fabsf(v0.w))),
                             fmaxf(fmaxf(fabsf(v1.x), fabsf(v1.y)), fmaxf(fabsf(v1.z), fabsf(v1.w)))));
        const __half2 h0 = __floats2half2_rn(v0.x, v0.y), h1 = __floats2half2_rn(v0.z, v0.w);
        const __half2 h2 = __floats2half2_rn(v1.x, v1.y), h3 = __floats2half2_rn(v1.z, v1.w);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
        const __half2 l0 = __floats2half2_rn(v0.x - f0.x, v0.y - f0.y), l1 = __floats2half2_rn(v0.z - f1.x, v0.w - f1.y);
        const __half2 l2 = __floats2half2_rn(v1.x - f2.x, v1.y - f2.y), l3 = __floats2half2_rn(v1.z - f3.x, v1.w - f3.y);
        uint4 hi, lo;
        hi.x = *reinterpret_cast<const uint32_t*>(&h0); hi.y = *reinterpret_cast<const uint32_t*>(&h1);
        hi.z = *reinterpret_cast<const uint32_t*>(&h2); hi.w = *reinterpret_cast<const uint32_t*>(&h3);
        lo.x = *reinterpret_cast<const uint32_t*>(&l0); lo.y = *reinterpret_cast<const uint32_t*>(&l1);
        lo.z = *reinterpret_cast<const uint32_t*>(&l2); lo.w = *reinterpret_cast<const uint32_t*>(&l3);
        const int kg = 4 * c32 + g4;                                // global K group of 8 elements
        const int chunk = kg / G, g = kg - chunk * G;
        const size_t o = ((((size_t)tile * nchunks + chunk) * 2) * G + g) * Rp + rl;
        dst[o] = hi;
        dst[o + (size_t)G * Rp] = lo;
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (g4 == 0) norms[(size_t)z * tiles * Rp + prow] = acc;
    mx = warp_max(mx);
    if (lane == 0 && !(mx < 60000.0f)) atomicOr(&ovf[z], 1u);
}

// Binary descriptors (hard-binarised BAD, descriptor/bad.py:560-574: every entry of a row is 0 or one common value s, the
// reciprocal of the row's norm, or 1 when the rows are not normalised): the similarity is s1 s2 popcount(bits1 & bits2), so the
// operands shrink to ONE 8-bit term (E4M3 1.0 / 0.0: a quarter of the bytes of the two fp16 terms, a third of the MMAs, exact
// integer accumulation).  out[z][tile][chunk][g][Rp rows][16 bytes]; squared norms as pack_f16_kernel; scales[row] = s; the
// per-pair flag is raised when some row is not of that form (the kernel then takes its FP32 dot-product path for the pair).
__global__ void __launch_bounds__(256) pack_bits_kernel(const float* d, int rows, int D, int Rp, int tiles, int G,
                                                        unsigned char* out, float* norms, float* scales, unsigned int* ovf) {
    const int z = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * 8 + (threadIdx.x >> 5);          // 8 consecutive padded rows
    const int prow = item * 8 + (lane >> 2);
    if (item * 8 >= tiles * Rp) return;
    const int g4 = lane & 3;                                       // 16 of every 64 elements
    const int tile = prow / Rp, rl = prow - tile * Rp;
    const int nchunks = D / (16 * G);
    const bool real = prow < rows;
    const float* src = d + ((size_t)z * rows + (real ? prow : 0)) * D + 16 * g4;
    uint4* dst = reinterpret_cast<uint4*>(out) + (size_t)z * tiles * nchunks * G * Rp;
    // one pass: bits, squared norm, the largest entry and the smallest non-zero entry of the row -- the row has the form
    // {0, s} exactly when the two agree (or the row is all zeros)
    float acc = 0.0f, mx = 0.0f, mnz = CUDART_INF_F;
    for (int c64 = 0; c64 < D / 64; ++c64) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = real ? __ldg(reinterpret_cast<const float4*>(src + 64 * c64 + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
            acc = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, acc))));
            mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
            mnz = fminf(mnz, fminf(fminf(v.x != 0.0f ? v.x : CUDART_INF_F, v.y != 0.0f ? v.y : CUDART_INF_F),
                                   fminf(v.z != 0.0f ? v.z : CUDART_INF_F, v.w != 0.0f ? v.w : CUDART_INF_F)));
            w[q] = (v.x != 0.0f ? 0x38u : 0u) | (v.y != 0.0f ? 0x3800u : 0u) | (v.z != 0.0f ? 0x380000u : 0u) | (v.w != 0.0f ? 0x38000000u : 0u);
        }
        const int kg = 4 * c64 + g4;                                // global K group of 16 elements
        const int chunk = kg / G, g = kg - chunk * G;
        dst[(((size_t)tile * nchunks + chunk) * G + g) * Rp + rl] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, 1));
    mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, 2));
    const bool odd = !(mnz == CUDART_INF_F || (mnz == mx && mx > 0.0f));   // a negative entry, two different values, NaN
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (g4 == 0) {
        norms[(size_t)z * tiles * Rp + prow] = acc;
        scales[(size_t)z * tiles * Rp + prow] = mx;
    }
    if (__any_sync(0xffffffffu, odd || !(mx < 60000.0f) || !(acc == acc)) && lane == 0) atomicOr(&ovf[z], 1u);
}

// out-of-fp16-range fallback of one similarity (rare): kept out of line, the unrolled epilogue calls it from 128 places
__device__ __noinline__ float dot_f32_slow(const float* x, const float* y, int D) {
    float dot = 0.0f;
    for (int k = 0; k < D; ++k) dot = fmaf(__ldg(x + k), __ldg(y + k), dot);
    return dot;
}

struct HyArgs {
    const unsigned char* d1p;   // packed fp16 terms of d1 / d2 (pack_f16_kernel)
    const unsigned char* d2p;
    const float* n1;            // squared norms, padded rows
    const float* n2;
    const unsigned int* ovf;    // per pair: some |x| >= 60000 (binary operands: some row is not {0, s}) -> FP32 dot products
    const float* s1;            // binary operands (pack_bits_kernel): per-row value of the set entries, padded rows; else null
    const float* s2;
    int binary;
    const float* d1;
    const float* d2;
    int N, M, D, Mp;
    int iterations;
    float scale2;               // log2(e)/eps
    float dustbin2;             // (-unused/eps) * log2(e)
    float* P;                   // EPI kernels: may be null (P never leaves the chip)
    long long* trace;           // optional (debug): per CTA 8 clock64 stamps
    SinkhornEpilogue e;
};

#define HY_STAMP(slot)                                                                                       \
    do {                                                                                                     \
        if (a.trace != nullptr && threadIdx.x == 0) a.trace[(size_t)blockIdx.x * 8 + (slot)] = clock64();    \
    } while (0)

template <int CL, int MAXM, int RRT, bool EPI>
__global__ void __launch_bounds__(NT, 1) sinkhorn_hy_kernel(HyArgs a) {
    using C = Hy<CL, MAXM, RRT>;
    constexpr int NK = C::NK, RR = C::RR, RS = C::RS, RPW = C::RPW, RPC = C::RPC, OWN = C::OWN, PSTR = C::PSTR, KC = C::KC, G = C::G;
    extern __shared__ __align__(128) float sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int z = blockIdx.x / CL;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = a.N, M = a.M, D = a.D, Mp = a.Mp;

    float* sKs = sm + C::OFF_KS;
    float* sCW = sm + C::OFF_CW;
    float* sB = sm + C::OFF_B;
    float* sPart = sm + C::OFF_PART;
    float* sAs = sm + C::OFF_AS;
    float* sN1 = sm + C::OFF_N1;
    float* sS1 = sm + C::OFF_S1;
    // [0,1] stage full (bulk copies landed), [2,3] stage free (MMAs retired), [4] GEMM done, [5] partials landed,
    // [6] b landed, [7] epilogue: sort keys landed (rank 0), [8,9] all-to-all partials landed (even / odd iterations)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int r0 = rank * RPC;
    const int nreal = max(0, min(RPC, N - r0));
    const bool has_dust = rank == CL - 1;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 10; ++i) mbar_init(smem_u32(&bars[i]), (i == 2 || i == 3) ? (uint32_t)CL : 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < RPC; i += NT) {                               // rows beyond N: +inf (their K becomes exactly 0)
        sN1[i] = r0 + i < N ? a.n1[((size_t)z * CL + rank) * RPC + i] : CUDART_INF_F;
        sS1[i] = (a.binary && r0 + i < N) ? a.s1[((size_t)z * CL + rank) * RPC + i] : 1.0f;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    cluster.sync();      // peers multicast into this CTA's staging area and arrive on its barriers from here on
    HY_STAMP(0);

    // ---------------- similarity GEMM on tcgen05: D[j][i] = sum_k d2[j][k] d1[i][k] ------------------
    const int nmb = Mp / 128;
    {
        // binary operands: one 8-bit term, 16 elements per 16-byte group (a stage holds 16 G elements instead of 8 G)
        const bool bin = a.binary != 0;
        const int nchunks = bin ? D / (16 * G) : D / KC;
        const uint32_t a_term = (uint32_t)G * Mp * 16u, b_term = (uint32_t)C::B_TERM;
        const uint32_t bytesA = (bin ? 1u : 2u) * a_term, bytesB = (bin ? 1u : 2u) * b_term;
        const uint32_t stage0 = smem_u32(sm);
        if (warp == 0 && lane == 0) {
            // ===== copy issuer: this CTA's d1 terms of the chunk, and ITS SLICE of the d2 terms multicast to the whole cluster
            // (every CTA needs all of d2: each of the CL issuers fetches 1 / CL of a chunk from L2 once for everybody) =====
            constexpr uint16_t ALL = (uint16_t)((1u << CL) - 1u);
            const size_t per_elem = bin ? 1 : 4;                       // packed bytes per descriptor element
            const unsigned char* gA = a.d2p + (size_t)z * (per_elem * D * Mp);
            const unsigned char* gB = a.d1p + ((size_t)z * CL + rank) * (per_elem * D * RPC);
            const uint32_t sliceA = bytesA / (uint32_t)CL;
            for (int c = 0; c < nchunks; ++c) {
                const int s = c & 1;
                // stage s is free in EVERY CTA: each one's MMA issuer commits chunk c-2 to all the cluster's barriers
                if (c >= 2) mbar_wait(smem_u32(&bars[2 + s]), (uint32_t)(((c >> 1) - 1) & 1));
                const uint32_t full = smem_u32(&bars[s]);
                mbar_arrive_expect_tx(full, bytesA + bytesB);
                bulk_g2s_mc(stage0 + s * C::STAGE + (uint32_t)rank * sliceA, gA + (size_t)c * bytesA + (size_t)rank * sliceA, sliceA,
                            full, ALL);
                bulk_g2s(stage0 + s * C::STAGE + 2 * C::A_TERM, gB + (size_t)c * bytesB, bytesB, full);
            }
        } else if (warp == 1 && lane == 0) {
            // ===== MMA issuer =====
            constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(RPC >> 3) << 17) | ((128u >> 4) << 24);   // D=F32, A=B=F16 (E4M3 under kind::f8f6f4), K-major, N=RPC, M=128
            const uint32_t a_lbo = (uint32_t)Mp * 16u, b_lbo = (uint32_t)RPC * 16u;
            for (int c = 0; c < nchunks; ++c) {
                const int s = c & 1;
                mbar_wait(smem_u32(&bars[s]), (uint32_t)((c >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = stage0 + s * C::STAGE;
                if (bin) {
#pragma unroll
                    for (int ks = 0; ks < G / 2; ++ks) {                // 32 elements = two 16-byte groups per instruction
                        const uint64_t bd = udesc(sa + 2 * C::A_TERM + 2 * ks * b_lbo, b_lbo, 128);
                        for (int mb = 0; mb < nmb; ++mb) {
                            const uint64_t ad = udesc(sa + 2 * ks * a_lbo + mb * 128 * 16, a_lbo, 128);
                            umma_f8_n(tmem_base + (uint32_t)(mb * RPC), ad, bd, IDESC, (c | ks) != 0);
                        }
                    }
                } else {
#pragma unroll
                    for (int ks = 0; ks < KC / 16; ++ks) {
                        const uint64_t bhi = udesc(sa + 2 * C::A_TERM + 2 * ks * b_lbo, b_lbo, 128);
                        const uint64_t blo = udesc(sa + 2 * C::A_TERM + b_term + 2 * ks * b_lbo, b_lbo, 128);
                        for (int mb = 0; mb < nmb; ++mb) {
                            const uint64_t ahi = udesc(sa + 2 * ks * a_lbo + mb * 128 * 16, a_lbo, 128);
                            const uint64_t alo = udesc(sa + a_term + 2 * ks * a_lbo + mb * 128 * 16, a_lbo, 128);
                            const uint32_t d = tmem_base + (uint32_t)(mb * RPC);
                            umma_f16_n(d, ahi, bhi, IDESC, (c | ks) != 0);
                            umma_f16_n(d, ahi, blo, IDESC, 1u);
                            umma_f16_n(d, alo, bhi, IDESC, 1u);
                        }
                    }
                }
                umma_commit_mc(smem_u32(&bars[2 + s]), (uint16_t)((1u << CL) - 1u));   // stage s of this CTA is free: tell every issuer
                if (c == nchunks - 1) umma_commit_to(smem_u32(&bars[4]));
            }
        }
        __syncwarp();
        HY_STAMP(1);
        mbar_wait(smem_u32(&bars[4]), 0u);                              // every MMA has completed
        HY_STAMP(2);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        __syncthreads();                                                // staging area is dead
    }

    // ---------------- epilogue: TMEM -> K = exp(-cost/eps) -> registers (first RR rows of a warp) / shared memory ----
    // warp (mbw, q) reads TMEM lanes 32q..32q+31 of M-blocks mbw, mbw+4: lane = column j, registers = 32 rows i.
    // A block of 32 rows belongs to 32/RPW consecutive warps; their register rows go through the bounce buffer.
    float4 kreg[RR][NK];
    {
        const bool overflow = a.ovf[z] != 0u;
        const float nscale2 = -a.scale2;
        const int mbw = warp >> 2, q = warp & 3;
        constexpr int WPB = 32 / RPW;                                   // warps per 32-row block
        for (int h = 0; h < RPC / 32; ++h) {
            for (int mb = mbw; mb < nmb; mb += 4) {
                const int j = 128 * mb + 32 * q + lane;
                uint32_t r[32];
                tmem_ld32x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(mb * RPC + 32 * h), r);
                const bool jv = j < M;
                // rows beyond N and columns beyond M carry a squared norm of +inf (sN1 / n2j): their cost is +inf and their
                // K is exactly 0 without a per-element test
                const float n2j = jv ? a.n2[(size_t)z * Mp + j] : CUDART_INF_F;
                // binary operands: the accumulator is popcount(bits_i & bits_j), the similarity s_i s_j popcount (s = 1 otherwise)
                const float m2s = -2.0f * ((a.binary && jv) ? a.s2[(size_t)z * Mp + j] : 1.0f);
                if (!overflow) {
#pragma unroll
                    for (int ii = 0; ii < 32; ++ii) {
                        const int li = 32 * h + ii;
                        // sinkhorn.py:98-103; 2 * dot is exact, so the fused form rounds exactly like (n1 + n2) - 2 * dot
                        const float cost = fmaxf(fmaf(m2s, __fmul_rn(__uint_as_float(r[ii]), sS1[li]), __fadd_rn(sN1[li], n2j)), 0.0f);
                        const float kv = ex2h(__fmul_rn(cost, nscale2));
                        const int wl = ii / RPW, ri = ii % RPW;         // compile-time after unrolling
                        if (ri >= RR) sKs[((WPB * h + wl) * RS + (ri - RR)) * MAXM + j] = kv;
                        else sCW[(wl * RR + ri) * MAXM + j] = kv;
                    }
                } else {                                                // out of fp16 range (rare): FP32 dot products, out of line
#pragma unroll
                    for (int ii = 0; ii < 32; ++ii) {
                        const int li = 32 * h + ii;
                        const bool ev = jv && li < nreal;
                        float kv = 0.0f;
                        if (ev) {
                            const float dot = dot_f32_slow(a.d1 + ((size_t)z * N + r0 + li) * D, a.d2 + ((size_t)z * M + j) * D, D);
                            const float cost = fmaxf(__fsub_rn(__fadd_rn(sN1[li], n2j), __fmul_rn(2.0f, dot)), 0.0f);
                            kv = ex2h(__fmul_rn(cost, nscale2));
                        }
                        const int wl = ii / RPW, ri = ii % RPW;
                        if (ri >= RR) sKs[((WPB * h + wl) * RS + (ri - RR)) * MAXM + j] = kv;
                        else sCW[(wl * RR + ri) * MAXM + j] = kv;
                    }
                }
            }
            if (nmb < C::NMB) {                                          // columns beyond Mp: zero
                for (int e2 = tid; e2 < 32 * (MAXM - Mp); e2 += NT) {
                    const int ii = e2 / (MAXM - Mp), j = Mp + e2 % (MAXM - Mp);
                    const int wl = ii / RPW, ri = ii % RPW;
                    if (ri >= RR) sKs[((WPB * h + wl) * RS + (ri - RR)) * MAXM + j] = 0.0f;
                    else sCW[(wl * RR + ri) * MAXM + j] = 0.0f;
                }
            }
            __syncthreads();
            if (warp / WPB == h) {
                const int wl = warp - WPB * h;
#pragma unroll
                for (int rr = 0; rr < RR; ++rr)
#pragma unroll
                    for (int k = 0; k < NK; ++k)
                        kreg[rr][k] = *reinterpret_cast<const float4*>(sCW + (wl * RR + rr) * MAXM + 128 * k + 4 * lane);
            }
            __syncthreads();
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                                // every accumulator has been read
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // The accumulators are dead: tensor memory becomes the home of the K rows that do not fit the registers.  A warp can
    // reach the 32 lanes of its quadrant (warp % 4); the four warps of a quadrant take 128 columns each, i.e. 128 private
    // 32-bit words per thread, read back at ~800 B/clk per SM (shared memory: 128 B/clk) -- measured, tools/probes/tmem_probe.cu.
    // Word (k * RS + sr) * 4 + q of a thread = K[row RR + sr][column 128 k + 4 lane + q].
    const uint32_t tpriv = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(128 * (warp >> 2));
    static_assert(RS * 4 * NK <= 128, "a thread's private tensor-memory words");
    {
        const float4* src = reinterpret_cast<const float4*>(sKs) + (size_t)warp * RS * (MAXM / 4) + lane;
#pragma unroll
        for (int k = 0; k < NK; ++k)
#pragma unroll
            for (int sr = 0; sr < RS; ++sr) tmem_st4(tpriv + (uint32_t)((k * RS + sr) * 4), src[sr * (MAXM / 4) + 32 * k]);
        tmem_wait_st();
    }
    HY_STAMP(3);

    // ---------------- Sinkhorn in scaling form (see sinkhorn_tc.cu) ------------------------------------------------
    // warp w owns rows RPW*w .. RPW*w + RPW-1 of this CTA (first RR in registers, last RR in shared memory), lane l
    // columns 128k + 4l .. +3.  Row sums stay inside the warp; column sums go warp -> CTA through shared memory and
    // CTA -> cluster as a reduce-scatter (CTA c finishes columns OWN*c .. OWN*c + OWN-1, the last rank also the dustbin
    // column) followed by an all-gather of the new b, both as st.async stores that complete on the receivers' mbarriers.
    const float kd = ex2h(a.dustbin2);
    const float Mf = (float)M, Nf = (float)N;
    for (int i = tid; i <= MAXM; i += NT) sB[i] = (i < M || i == MAXM) ? 1.0f : 0.0f;
    const uint32_t my_part = smem_u32(sPart + rank * PSTR), loc_b = smem_u32(sB);
    const uint32_t loc_bar_part = smem_u32(&bars[5]), loc_bar_b = smem_u32(&bars[6]);
    const uint32_t part_bytes = (uint32_t)CL * (has_dust ? (uint32_t)OWN + 1u : (uint32_t)OWN) * 4u;
    const uint32_t b_bytes = (uint32_t)(MAXM + 1) * 4u;
    const float4* sB4 = reinterpret_cast<const float4*>(sB);
    float4* sCW4 = reinterpret_cast<float4*>(sCW) + (size_t)warp * (MAXM / 4) + lane;
    cluster.sync();      // every CTA is past its GEMM and has initialised its barriers and its b
    HY_STAMP(4);

    float av[RPW];
    float aN = 0.0f, bM = 1.0f;
#pragma unroll
    for (int r = 0; r < RPW; ++r) av[r] = 0.0f;
    float* sX = sm + C::OFF_X;
    for (int it = 0; it < a.iterations; ++it) {
        const uint32_t par = (uint32_t)(it & 1);
        const uint32_t xbuf = (uint32_t)(it & 1), xpar = (uint32_t)((it >> 1) & 1);
        const uint32_t loc_bar_x = smem_u32(&bars[8 + xbuf]);
        if (tid == 0) {
            if constexpr (C::A2A) {
                mbar_arrive_expect_tx(loc_bar_x, (uint32_t)CL * (uint32_t)(MAXM + 1) * 4u);
            } else {
                mbar_arrive_expect_tx(loc_bar_part, part_bytes);
                mbar_arrive_expect_tx(loc_bar_b, b_bytes);
            }
        }
        // ---- a_i = mu_i / rowsum_i ----
        float rs[RPW], sb = 0.0f;
        float2 rs2[RPW];
#pragma unroll
        for (int r = 0; r < RPW; ++r) rs2[r] = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const float4 b4 = sB4[32 * k + lane];
            sb += (b4.x + b4.y) + (b4.z + b4.w);
            float4 kt[RS];
#pragma unroll
            for (int sr = 0; sr < RS; ++sr) tmem_ld4(tpriv + (uint32_t)((k * RS + sr) * 4), kt[sr]);
#pragma unroll
            for (int rr = 0; rr < RR; ++rr) rs2[rr] = dot4x2(kreg[rr][k], b4, rs2[rr]);      // while the tensor-memory loads land
            tmem_wait_ld();
#pragma unroll
            for (int sr = 0; sr < RS; ++sr) rs2[RR + sr] = dot4x2(kt[sr], b4, rs2[RR + sr]);
        }
#pragma unroll
        for (int r = 0; r < RPW; ++r) rs[r] = rs2[r].x + rs2[r].y;
        bM = sB[MAXM];
        float asum = 0.0f;
        if constexpr (RPW == 8) {
            // transposing butterfly: after the steps with partners 16, 8 and 4 a lane carries ONE row's partial sum (row
            // lane >> 2), so the five steps take 4 + 2 + 1 + 1 + 1 shuffles instead of 8 x 5; the additions of a row are those
            // of the plain xor butterfly (16, 8, 4, 2, 1) in the same order
            float x[4], y[2], zr;
            {
                const bool hi = (lane & 16) != 0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float send = hi ? rs[r] : rs[r + 4], keep = hi ? rs[r + 4] : rs[r];
                    x[r] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
            }
            {
                const bool hi = (lane & 8) != 0;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const float send = hi ? x[r] : x[r + 2], keep = hi ? x[r + 2] : x[r];
                    y[r] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
            }
            {
                const bool hi = (lane & 4) != 0;
                const float send = hi ? y[0] : y[1], keep = hi ? y[1] : y[0];
                zr = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            zr += __shfl_xor_sync(0xffffffffu, zr, 2);
            zr += __shfl_xor_sync(0xffffffffu, zr, 1);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sb += __shfl_xor_sync(0xffffffffu, sb, o);
            const float mine = (RPW * warp + (lane >> 2) < nreal) ? __fdividef(1.0f, fmaf(kd, bM, zr)) : 0.0f;
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                av[r] = __shfl_sync(0xffffffffu, mine, 4 * r);
                asum += av[r];
            }
        } else {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int r = 0; r < RPW; ++r) rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], o);
                sb += __shfl_xor_sync(0xffffffffu, sb, o);
            }
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                av[r] = (RPW * warp + r < nreal) ? __fdividef(1.0f, fmaf(kd, bM, rs[r])) : 0.0f;
                asum += av[r];
            }
        }
        aN = __fdividef(Mf, kd * (sb + bM));                            // dustbin row: mu_N = M (sinkhorn.py:197-198)
        // ---- column sums: warp partials -> CTA partials -> owners ----
        float2 av2[RPW];
#pragma unroll
        for (int r = 0; r < RPW; ++r) av2[r] = make_float2(av[r], av[r]);
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            float4 kt[RS];
#pragma unroll
            for (int sr = 0; sr < RS; ++sr) tmem_ld4(tpriv + (uint32_t)((k * RS + sr) * 4), kt[sr]);
            float2 lo = make_float2(0.0f, 0.0f), hi = lo;
#pragma unroll
            for (int rr = 0; rr < RR; ++rr) axpy4x2(kreg[rr][k], av2[rr], lo, hi);
            tmem_wait_ld();
#pragma unroll
            for (int sr = 0; sr < RS; ++sr) axpy4x2(kt[sr], av2[RR + sr], lo, hi);
            sCW4[32 * k] = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
        if (lane == 0) sAs[warp] = asum;
        bar_all();
        if constexpr (C::A2A) {
            // ---- one hop: every CTA sends its partials of ALL columns to every CTA (itself included, so that one mbarrier
            // covers the lot), then each CTA finishes every b_j itself: same additions in the same order in all CTAs ----
            if (tid < MAXM / 4) {                                       // four adjacent columns per thread
                const float4* c4 = reinterpret_cast<const float4*>(sCW) + tid;
                float4 s0 = c4[0], s1 = c4[MAXM / 4], s2 = c4[2 * (MAXM / 4)], s3 = c4[3 * (MAXM / 4)];
#pragma unroll
                for (int w = 4; w < NW; w += 4) {
                    const float4 q0 = c4[w * (MAXM / 4)], q1 = c4[(w + 1) * (MAXM / 4)], q2 = c4[(w + 2) * (MAXM / 4)],
                                 q3 = c4[(w + 3) * (MAXM / 4)];
                    s0.x += q0.x; s0.y += q0.y; s0.z += q0.z; s0.w += q0.w;
                    s1.x += q1.x; s1.y += q1.y; s1.z += q1.z; s1.w += q1.w;
                    s2.x += q2.x; s2.y += q2.y; s2.z += q2.z; s2.w += q2.w;
                    s3.x += q3.x; s3.y += q3.y; s3.z += q3.z; s3.w += q3.w;
                }
                float4 sum;
                sum.x = (s0.x + s1.x) + (s2.x + s3.x); sum.y = (s0.y + s1.y) + (s2.y + s3.y);
                sum.z = (s0.z + s1.z) + (s2.z + s3.z); sum.w = (s0.w + s1.w) + (s2.w + s3.w);
                const uint32_t dst = smem_u32(sX + ((int)xbuf * CL + rank) * C::XSTR + 4 * tid);
#pragma unroll
                for (int p = 0; p < CL; ++p) st_async_v4(mapa32(dst, (uint32_t)p), sum, mapa32(loc_bar_x, (uint32_t)p));
            } else if (tid == MAXM / 4) {
                float as = 0.0f;
#pragma unroll
                for (int w = 0; w < NW; ++w) as += sAs[w];
                const uint32_t dst = smem_u32(sX + ((int)xbuf * CL + rank) * C::XSTR + MAXM);
#pragma unroll
                for (int p = 0; p < CL; ++p) st_async32(mapa32(dst, (uint32_t)p), kd * as, mapa32(loc_bar_x, (uint32_t)p));
            }
            mbar_wait(loc_bar_x, xpar);
            const float* xb = sX + (int)xbuf * CL * C::XSTR;
#pragma unroll
            for (int cc = 0; cc < MAXM / NT; ++cc) {
                const int col = tid + NT * cc;
                float t = 0.0f;
#pragma unroll
                for (int r = 0; r < CL; ++r) t += xb[r * C::XSTR + col];
                t = fmaf(kd, aN, t);
                sB[col] = col < M ? __fdividef(1.0f, t) : 0.0f;
            }
            if (tid == NT - 1) {
                float t = 0.0f;
#pragma unroll
                for (int r = 0; r < CL; ++r) t += xb[r * C::XSTR + MAXM];
                t = fmaf(kd, aN, t);
                sB[MAXM] = __fdividef(Nf, t);                           // nu_M = N (:199-200)
            }
            bar_all();
        } else {
#pragma unroll
            for (int cc = 0; cc < MAXM / NT; ++cc) {                    // one column per thread (two at MAXM = 1024)
                const int col = tid + NT * cc;
                float s0 = sCW[col], s1 = sCW[MAXM + col], s2 = sCW[2 * MAXM + col], s3 = sCW[3 * MAXM + col];
#pragma unroll
                for (int w = 4; w < NW; w += 4) {
                    s0 += sCW[w * MAXM + col]; s1 += sCW[(w + 1) * MAXM + col];
                    s2 += sCW[(w + 2) * MAXM + col]; s3 += sCW[(w + 3) * MAXM + col];
                }
                const uint32_t owner = (uint32_t)(col / OWN);
                st_async32(mapa32(my_part + (uint32_t)(col % OWN) * 4u, owner), (s0 + s1) + (s2 + s3), mapa32(loc_bar_part, owner));
            }
            if (tid == 0) {
                float as = 0.0f;
#pragma unroll
                for (int w = 0; w < NW; ++w) as += sAs[w];
                st_async32(mapa32(my_part + (uint32_t)OWN * 4u, (uint32_t)(CL - 1)), kd * as, mapa32(loc_bar_part, (uint32_t)(CL - 1)));
            }
            // ---- owners: b_j = nu_j / colsum_j, sent to every CTA ----
            mbar_wait(loc_bar_part, par);
            if (tid < OWN || (tid == OWN && has_dust)) {
                float t = 0.0f;
#pragma unroll
                for (int r = 0; r < CL; ++r) t += sPart[r * PSTR + tid];
                const int c = tid == OWN ? M : OWN * rank + tid;        // global column (M = dustbin column)
                t = fmaf(kd, aN, t);
                const float bnew = (tid == OWN) ? __fdividef(Nf, t) : (c < M ? __fdividef(1.0f, t) : 0.0f);   // nu_M = N (:199-200)
                const uint32_t slot = (uint32_t)(tid == OWN ? MAXM : c) * 4u;
#pragma unroll
                for (int dst = 0; dst < CL; ++dst)
                    st_async32(mapa32(loc_b + slot, (uint32_t)dst), bnew, mapa32(loc_bar_b, (uint32_t)dst));
            }
            mbar_wait(loc_bar_b, par);
        }
    }
    bM = sB[MAXM];
    HY_STAMP(5);

    // ---------------- P = a_i K_ij b_j (sinkhorn.py:145, :206) -------------------------------------------------------
    float* Pz = a.P != nullptr ? a.P + (size_t)z * (N + 1) * (M + 1) : nullptr;
    if constexpr (!EPI) {
        float* stage = sCW + warp * MAXM;                               // this warp's slice: transposes to coalesced stores
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int li = RPW * warp + r;
            if (li < nreal) {                                           // warp-uniform
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    const float4 b4 = sB4[32 * k + lane];
                    float4 kv;
                    if (r < RR) kv = kreg[r < RR ? r : 0][k];
                    else { tmem_ld4(tpriv + (uint32_t)((k * RS + (r >= RR ? r - RR : 0)) * 4), kv); tmem_wait_ld(); }
                    float4 p4;
                    p4.x = av[r] * kv.x * b4.x; p4.y = av[r] * kv.y * b4.y; p4.z = av[r] * kv.z * b4.z; p4.w = av[r] * kv.w * b4.w;
                    *reinterpret_cast<float4*>(stage + 128 * k + 4 * lane) = p4;
                }
                __syncwarp();
                float* out = Pz + (size_t)(r0 + li) * (M + 1);
#pragma unroll
                for (int k = 0; k < MAXM / 32; ++k) {
                    const int c = lane + 32 * k;
                    if (c < M) out[c] = stage[c];
                }
                if (lane == 0) out[M] = av[r] * kd * bM;
                __syncwarp();
            }
        }
        if (has_dust && warp == 0) {                                    // dustbin row
            float* out = Pz + (size_t)N * (M + 1);
            const float f = aN * kd;
            for (int c = lane; c < M; c += 32) out[c] = f * sB[c];
            if (lane == 0) out[M] = f * bM;
        }
    } else {
        // ======== fused epilogue: filters (sinkhorn.py:311-465), scores (:211-259), mutual matches (match_extraction.py:46-184)
        // on the P values while they are on the chip.  P itself is written only when the caller wants it.
        const SinkhornEpilogue& e = a.e;
        const uint32_t epar = C::A2A ? 0u : (uint32_t)(a.iterations & 1);   // phase of bars[5] / bars[6] after the loop
        // CTA-local tables reuse b and the partial table (idle once the iterations are over and P has been formed); the
        // tables that PEERS write into have their own space: a faster peer may send while this CTA still forms its P rows
        float* eColBest = sB;                                           // [MAXM] CTA-level column maxima
        int* eColRow = reinterpret_cast<int*>(sPart);                   // [MAXM] first local row attaining them
        int* eColArg = reinterpret_cast<int*>(sm + C::OFF_MISC);        // [MAXM] argmax row of every column (all-gathered)
        unsigned long long* ePart64 = reinterpret_cast<unsigned long long*>(sm + C::OFF_MISC + MAXM);   // [CL][OWN]
        unsigned long long* eKeys = reinterpret_cast<unsigned long long*>(sm + C::OFF_MISC + 3 * MAXM); // [CL*RPC] (rank 0)
        const bool want_cols = e.scores1 != nullptr || e.matches;
        float* stage = sCW + warp * MAXM;
        float rbest[RPW];
        int rarg[RPW];
        // ---- P in place of K, row statistics, filters, optional P write; column maxima of this warp's rows ----
        float4 cbest[NK];                                               // best value per owned lane column over the warp's rows
        int4 crow[NK];
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            cbest[k] = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
            crow[k] = make_int4(0, 0, 0, 0);
        }
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int li = RPW * warp + r;
            float4 p[NK];
            float b1 = -CUDART_INF_F, b2 = -CUDART_INF_F;
            int arg = 0x7fffffff;
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                const float4 b4 = sB4[32 * k + lane];
                float4 kv;
                if (r < RR) kv = kreg[r < RR ? r : 0][k];
                else { tmem_ld4(tpriv + (uint32_t)((k * RS + (r >= RR ? r - RR : 0)) * 4), kv); tmem_wait_ld(); }
                p[k].x = av[r] * kv.x * b4.x; p[k].y = av[r] * kv.y * b4.y; p[k].z = av[r] * kv.z * b4.z; p[k].w = av[r] * kv.w * b4.w;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int col = 128 * k + 4 * lane + q;
                    const float v = q == 0 ? p[k].x : (q == 1 ? p[k].y : (q == 2 ? p[k].z : p[k].w));
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
            float pdust = av[r] * kd * bM;
            bool ok = true;
            if (e.filters) {
                if (e.ratio_threshold > 0.0f) {
                    const float second = M >= 2 ? b2 : 0.0f;            // sinkhorn.py:339-341
                    ok = ok && (__fdiv_rn(b1, __fadd_rn(second, 1e-8f)) >= e.ratio_threshold);
                }
                if (e.dustbin_margin >= 0.0f) ok = ok && (__fsub_rn(b1, pdust) >= e.dustbin_margin);
                if (!ok) {                                              // rejected: core -> 0, dustbin -> 1 (:448-457)
#pragma unroll
                    for (int k = 0; k < NK; ++k) p[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    pdust = 1.0f;
                    b1 = 0.0f;
                    arg = 0;
                }
                if (lane == 0 && li < nreal && e.filter_valid != nullptr) e.filter_valid[(size_t)z * N + r0 + li] = ok ? 1 : 0;
            }
            rbest[r] = b1;
            rarg[r] = arg;
            if (lane == 0 && li < nreal && e.scores0 != nullptr) e.scores0[(size_t)z * N + r0 + li] = b1;
            if (li < nreal) {                                           // ascending rows, strict >: the first maximal row wins
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    if (p[k].x > cbest[k].x) { cbest[k].x = p[k].x; crow[k].x = li; }
                    if (p[k].y > cbest[k].y) { cbest[k].y = p[k].y; crow[k].y = li; }
                    if (p[k].z > cbest[k].z) { cbest[k].z = p[k].z; crow[k].z = li; }
                    if (p[k].w > cbest[k].w) { cbest[k].w = p[k].w; crow[k].w = li; }
                }
            }
            if (Pz != nullptr && li < nreal) {                          // warp-uniform
#pragma unroll
                for (int k = 0; k < NK; ++k) *reinterpret_cast<float4*>(stage + 128 * k + 4 * lane) = p[k];
                __syncwarp();
                float* out = Pz + (size_t)(r0 + li) * (M + 1);
#pragma unroll
                for (int k = 0; k < MAXM / 32; ++k) {
                    const int c = lane + 32 * k;
                    if (c < M) out[c] = stage[c];
                }
                if (lane == 0) out[M] = pdust;
                __syncwarp();
            }
        }
        if (Pz != nullptr && has_dust && warp == 0) {                   // dustbin row: never filtered (:459-460)
            float* out = Pz + (size_t)N * (M + 1);
            const float f = aN * kd;
            for (int c = lane; c < M; c += 32) out[c] = f * sB[c];
            if (lane == 0) out[M] = f * bM;
        }
        if (want_cols) {
            if (tid == 0) {
                mbar_arrive_expect_tx(loc_bar_part, (uint32_t)CL * (uint32_t)OWN * 8u);
                if (e.matches) {
                    mbar_arrive_expect_tx(loc_bar_b, (uint32_t)MAXM * 4u);
                    if (rank == 0) mbar_arrive_expect_tx(smem_u32(&bars[7]), (uint32_t)(CL * RPC) * 8u);
                }
            }
            // ---- column statistics: warp partial -> CTA (value first, then the first row attaining it) -> owner ----
            bar_all();                                                  // every warp is done with its staging slice and with b
#pragma unroll
            for (int k = 0; k < NK; ++k) sCW4[32 * k] = cbest[k];
            for (int i = tid; i < MAXM; i += NT) eColRow[i] = 0x7fffffff;
            bar_all();
#pragma unroll
            for (int cc = 0; cc < MAXM / NT; ++cc) {
                const int col = tid + NT * cc;
                float best = sCW[col];
#pragma unroll
                for (int w = 1; w < NW; ++w) best = fmaxf(best, sCW[w * MAXM + col]);
                eColBest[col] = best;
            }
            bar_all();
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                const float4 cb = *reinterpret_cast<const float4*>(eColBest + 128 * k + 4 * lane);
                int* cr = eColRow + 128 * k + 4 * lane;
                if (cbest[k].x == cb.x && cb.x > -CUDART_INF_F) atomicMin(cr + 0, crow[k].x);
                if (cbest[k].y == cb.y && cb.y > -CUDART_INF_F) atomicMin(cr + 1, crow[k].y);
                if (cbest[k].z == cb.z && cb.z > -CUDART_INF_F) atomicMin(cr + 2, crow[k].z);
                if (cbest[k].w == cb.w && cb.w > -CUDART_INF_F) atomicMin(cr + 3, crow[k].w);
            }
            bar_all();
            // CTA result of every column -> its owner: one 64-bit word (value bits, global row)
#pragma unroll
            for (int cc = 0; cc < MAXM / NT; ++cc) {
                const int col = tid + NT * cc;
                const float best = eColBest[col];
                const int row = eColRow[col];
                const unsigned long long w64 = ((unsigned long long)__float_as_uint(best) << 32) |
                                               (unsigned long long)(unsigned)(row == 0x7fffffff ? 0 : r0 + row);
                const uint32_t owner = (uint32_t)(col / OWN);
                st_async64(mapa32(smem_u32(ePart64 + rank * OWN + (col % OWN)), owner), w64, mapa32(loc_bar_part, owner));
            }
            mbar_wait(loc_bar_part, epar);
            if (tid < OWN) {                                            // owner: best over the row slices, lowest rank first
                unsigned long long w0 = ePart64[tid];
                float best = __uint_as_float((unsigned)(w0 >> 32));
                int arg = (int)(unsigned)(w0 & 0xffffffffu);
#pragma unroll
                for (int r = 1; r < CL; ++r) {
                    const unsigned long long w1 = ePart64[r * OWN + tid];
                    const float v = __uint_as_float((unsigned)(w1 >> 32));
                    if (v > best) { best = v; arg = (int)(unsigned)(w1 & 0xffffffffu); }
                }
                const int col = OWN * rank + tid;
                if (col < M && e.scores1 != nullptr) e.scores1[(size_t)z * M + col] = best;
                if (e.matches) {
#pragma unroll
                    for (int dst = 0; dst < CL; ++dst)
                        st_async32(mapa32(smem_u32(eColArg + col), (uint32_t)dst), __int_as_float(arg), mapa32(loc_bar_b, (uint32_t)dst));
                }
            }
        }
        if (e.matches) {
            mbar_wait(loc_bar_b, epar);
            // ---- mutual check + threshold (match_extraction.py:95-121): one 64-bit key per row to rank 0 ----
            if (lane < RPW) {
                float b1 = 0.0f;
                int arg = 0;
#pragma unroll
                for (int r = 0; r < RPW; ++r)
                    if (lane == r) { b1 = rbest[r]; arg = rarg[r]; }
                const int li = RPW * warp + lane;
                unsigned long long key = 0ull;                          // rows beyond N sort last
                if (li < nreal) {
                    const int gi = r0 + li;
                    const bool mutual = eColArg[arg] == gi;
                    const float sc = (mutual && b1 >= e.match_threshold) ? b1 : -1.0f;
                    key = ((unsigned long long)ordered_bits_hy(sc) << 32) |
                          (unsigned long long)(((0xFFFFu - (unsigned)gi) << 16) | (unsigned)arg);
                }
                st_async64(mapa32(smem_u32(eKeys + (r0 + li)), 0u), key, mapa32(smem_u32(&bars[7]), 0u));
            }
            if (rank == 0) {
                constexpr int NKEY = CL * RPC;
                mbar_wait(smem_u32(&bars[7]), 0u);
                const unsigned long long key0 = eKeys[0];               // row 0's key (its argmax column pads short outputs)
                bar_all();
                // descending bitonic sort of the keys (torch.topk sorted, :124-130; equal scores in ascending row order)
                for (int size = 2; size <= NKEY; size <<= 1) {
                    for (int stride = size >> 1; stride > 0; stride >>= 1) {
                        for (int t = tid; t < NKEY / 2; t += NT) {
                            const int lo = 2 * t - (t & (stride - 1));
                            const int hi = lo + stride;
                            const bool desc = (lo & size) == 0;
                            const unsigned long long x = eKeys[lo], y = eKeys[hi];
                            if ((x < y) == desc) { eKeys[lo] = y; eKeys[hi] = x; }
                        }
                        bar_all();
                    }
                }
                const int take = min(e.max_matches, N);
                for (int m = tid; m < e.max_matches; m += NT) {
                    float sc = 0.0f;                                    // zero padding when N < max_matches (:133-142)
                    unsigned long long key = key0;
                    if (m < take) {
                        key = eKeys[m];
                        sc = from_ordered_bits_hy((unsigned int)(key >> 32));
                    }
                    const int i = m < take ? (int)(0xFFFFu - (unsigned)((key >> 16) & 0xFFFFu)) : 0;
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
    }
    HY_STAMP(6);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster.sync();      // no CTA may exit while a peer can still write into its shared memory; every warp is done with tensor memory
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
}

template <int CL, int MAXM, int RRT, bool EPI>
int launch_hy(const HyArgs& a, int B, cudaStream_t st) {
    using C = Hy<CL, MAXM, RRT>;
    auto kern = sinkhorn_hy_kernel<CL, MAXM, RRT, EPI>;
    const size_t smem = (size_t)C::SMEM_FLOATS * sizeof(float);
    OM_TRY(set_smem(kern, smem));
    if (CL > 8) OM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(B * CL), 1, 1);
    cfg.blockDim = dim3(NT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    OM_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    OM_AFTER_LAUNCH();
    return OM_OK;
}

}  // namespace

// the packing kernel for other callers (sinkhorn_xl.cu): `tiles` tiles of Rp rows per pair, G 16-byte K groups per chunk
int pack_f16_launch(const float* d, int B, int rows, int D, int Rp, int tiles, int G, unsigned char* out, float* norms,
                    unsigned int* ovf, cudaStream_t st) {
    const size_t np = (size_t)tiles * Rp;
    pack_f16_kernel<<<dim3((unsigned)((np / 8 + 7) / 8), (unsigned)B), 256, 0, st>>>(d, rows, D, Rp, tiles, G, out, norms, ovf);
    OM_AFTER_LAUNCH();
    return OM_OK;
}

extern long long* g_tc_trace;       // sinkhorn_tc.cu (debug stamps)
int g_hy_allow_16 = 1;              // test hook: 0 keeps K in (512, 1024] on the generic path

// limits of the hybrid kernel: squared-L2 cost, scaling form in range, D % 32 == 0, N, M <= 1024
bool sinkhorn_hy_eligible(int N, int M, int D, float eps, float unused, int distance_l1) {
    const double log2e = 1.4426950408889634;
    if (distance_l1 || D % 32 != 0 || D <= 0 || N <= 0 || M <= 0) return false;
    if (!((double)unused / (double)eps * log2e <= 60.0) || !(unused >= 0.0f)) return false;
    if (N <= 512 && M <= 512) return true;
    return g_hy_allow_16 && N <= 1024 && M <= 1024;
}

static void hy_geometry(int N, int M, int& CLv, int& RPCv, int& Mp) {
    const bool big = N > 512 || M > 512;
    CLv = big ? 16 : 4;
    RPCv = big ? 64 : 128;
    Mp = (M + 127) / 128 * 128;
}

size_t sinkhorn_hy_workspace_bytes(int B, int N, int M, int D) {
    if (B <= 0 || N <= 0 || M <= 0 || D <= 0 || N > 1024 || M > 1024) return 0;
    int CLv, RPCv, Mp;
    hy_geometry(N, M, CLv, RPCv, Mp);
    const size_t np = (size_t)CLv * RPCv;
    return align_up((size_t)B * np * D * 4) + align_up((size_t)B * Mp * D * 4) + 2 * align_up((size_t)B * np * sizeof(float)) +
           2 * align_up((size_t)B * Mp * sizeof(float)) + align_up((size_t)B * sizeof(unsigned int));
}

namespace {
struct HyWs {
    unsigned char* d1p;
    unsigned char* d2p;
    float* n1;
    float* n2;
    float* s1;                  // binary operands: per-row values
    float* s2;
    unsigned int* ovf;
    int CLv, RPCv, Mp, G;
};
HyWs hy_carve(void* ws, int B, int N, int M, int D) {
    HyWs w{};
    hy_geometry(N, M, w.CLv, w.RPCv, w.Mp);
    const size_t np = (size_t)w.CLv * w.RPCv;
    char* c = (char*)ws;
    w.d1p = (unsigned char*)c; c += align_up((size_t)B * np * D * 4);
    w.d2p = (unsigned char*)c; c += align_up((size_t)B * w.Mp * D * 4);
    w.n1 = (float*)c; c += align_up((size_t)B * np * sizeof(float));
    w.n2 = (float*)c; c += align_up((size_t)B * w.Mp * sizeof(float));
    w.s1 = (float*)c; c += align_up((size_t)B * np * sizeof(float));
    w.s2 = (float*)c; c += align_up((size_t)B * w.Mp * sizeof(float));
    w.ovf = (unsigned int*)c;
    w.G = w.CLv == 4 ? 4 : 2;
    return w;
}
}  // namespace

// The three steps of a hybrid-kernel call, separately launchable so that the fused matcher can pack each image's descriptors
// on that image's own stream as soon as they exist (the packing of one image then runs under the other image's kernels):
// prepare (clears the per-pair range flags), pack (which = 0: desc1 rows, 1: desc2 rows), run (the cluster kernel).
int sinkhorn_hy_prepare(int B, int N, int M, int D, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (ws == nullptr || ws_bytes < sinkhorn_hy_workspace_bytes(B, N, M, D)) return OM_ERR_WORKSPACE;
    const HyWs w = hy_carve(ws, B, N, M, D);
    OM_CUDA(cudaMemsetAsync(w.ovf, 0, (size_t)B * sizeof(unsigned int), st));
    return OM_OK;
}

// binary operands need whole 64-element groups per lane of the packing kernel and whole stages of 16 G elements
bool sinkhorn_hy_binary_ok(int N, int M, int D) {
    int CLv, RPCv, Mp;
    hy_geometry(N, M, CLv, RPCv, Mp);
    const int G = CLv == 4 ? 4 : 2;
    return D % 64 == 0 && D % (16 * G) == 0;
}

int sinkhorn_hy_pack(int which, const float* d, int B, int N, int M, int D, void* ws, size_t ws_bytes, cudaStream_t st, int binary) {
    if (d == nullptr) return OM_ERR_NULL;
    if (ws == nullptr || ws_bytes < sinkhorn_hy_workspace_bytes(B, N, M, D)) return OM_ERR_WORKSPACE;
    if (binary && !sinkhorn_hy_binary_ok(N, M, D)) return OM_ERR_PARAM;
    const HyWs w = hy_carve(ws, B, N, M, D);
    const size_t np = (size_t)w.CLv * w.RPCv;
    if (binary) {
        if (which == 0)
            pack_bits_kernel<<<dim3((unsigned)((np / 8 + 7) / 8), (unsigned)B), 256, 0, st>>>(d, N, D, w.RPCv, w.CLv, w.G, w.d1p, w.n1, w.s1, w.ovf);
        else
            pack_bits_kernel<<<dim3((unsigned)((w.Mp / 8 + 7) / 8), (unsigned)B), 256, 0, st>>>(d, M, D, w.Mp, 1, w.G, w.d2p, w.n2, w.s2, w.ovf);
        OM_AFTER_LAUNCH();
        return OM_OK;
    }
    if (which == 0)
        pack_f16_kernel<<<dim3((unsigned)((np / 8 + 7) / 8), (unsigned)B), 256, 0, st>>>(d, N, D, w.RPCv, w.CLv, w.G, w.d1p, w.n1, w.ovf);
    else
        pack_f16_kernel<<<dim3((unsigned)((w.Mp / 8 + 7) / 8), (unsigned)B), 256, 0, st>>>(d, M, D, w.Mp, 1, w.G, w.d2p, w.n2, w.ovf);
    OM_AFTER_LAUNCH();
    return OM_OK;
}

int sinkhorn_hy_run(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float eps, float unused,
                    float* P, const SinkhornEpilogue* e, void* ws, size_t ws_bytes, cudaStream_t st, int binary) {
    if (!sinkhorn_hy_eligible(N, M, D, eps, unused, 0)) return OM_ERR_PARAM;
    if (ws == nullptr || ws_bytes < sinkhorn_hy_workspace_bytes(B, N, M, D)) return OM_ERR_WORKSPACE;
    if (binary && !sinkhorn_hy_binary_ok(N, M, D)) return OM_ERR_PARAM;
    const HyWs w = hy_carve(ws, B, N, M, D);
    HyArgs a{};
    a.d1p = w.d1p; a.d2p = w.d2p; a.n1 = w.n1; a.n2 = w.n2; a.ovf = w.ovf; a.d1 = d1; a.d2 = d2;
    a.binary = binary ? 1 : 0; a.s1 = w.s1; a.s2 = w.s2;
    a.N = N; a.M = M; a.D = D; a.Mp = w.Mp; a.iterations = iterations; a.P = P;
    a.trace = g_tc_trace;
    const double log2e = 1.4426950408889634;
    a.scale2 = (float)(log2e / (double)eps);
    a.dustbin2 = (float)((-(double)unused / (double)eps) * log2e);
    const bool epi = e != nullptr && e->any();
    if (epi) a.e = *e;
    if (w.CLv == 4) return epi ? launch_hy<4, 512, HY_RR4, true>(a, B, st) : launch_hy<4, 512, HY_RR4, false>(a, B, st);
    return epi ? launch_hy<16, 1024, 2, true>(a, B, st) : launch_hy<16, 1024, 2, false>(a, B, st);
}

int sinkhorn_hy_launch(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float eps, float unused,
                       float* P, const SinkhornEpilogue* e, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!sinkhorn_hy_eligible(N, M, D, eps, unused, 0)) return OM_ERR_PARAM;
    OM_TRY(sinkhorn_hy_prepare(B, N, M, D, ws, ws_bytes, st));
    OM_TRY(sinkhorn_hy_pack(0, d1, B, N, M, D, ws, ws_bytes, st));
    OM_TRY(sinkhorn_hy_pack(1, d2, B, N, M, D, ws, ws_bytes, st));
    return sinkhorn_hy_run(d1, d2, B, N, M, D, iterations, eps, unused, P, e, ws, ws_bytes, st);
}

// debug: how many clusters of the hybrid kernel the device can hold at once (cudaOccupancyMaxActiveClusters)
int hy_max_active_clusters(int big) {
    int n = -1;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cfg.blockDim = dim3(NT, 1, 1);
    if (big) {
        using C = Hy<16, 1024, 2>;
        auto kern = sinkhorn_hy_kernel<16, 1024, 2, false>;
        cfg.dynamicSmemBytes = (size_t)C::SMEM_FLOATS * sizeof(float);
        if (set_smem(kern, cfg.dynamicSmemBytes) != OM_OK) return -2;
        cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cfg.gridDim = dim3(16 * 64, 1, 1);
        at[0].val.clusterDim.x = 16;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return -3; }
    } else {
        using C = Hy<4, 512, HY_RR4>;
        auto kern = sinkhorn_hy_kernel<4, 512, HY_RR4, false>;
        cfg.dynamicSmemBytes = (size_t)C::SMEM_FLOATS * sizeof(float);
        if (set_smem(kern, cfg.dynamicSmemBytes) != OM_OK) return -2;
        cfg.gridDim = dim3(4 * 64, 1, 1);
        at[0].val.clusterDim.x = 4;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return -3; }
    }
    return n;
}

}  // namespace om

extern "C" int om_debug_hy_max_clusters(int big) { return om::hy_max_active_clusters(big); }
