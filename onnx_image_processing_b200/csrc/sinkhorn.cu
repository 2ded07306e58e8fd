// Matching stage: descriptor cost matrix + log-domain Sinkhorn with dustbins.
//
// Replaces matching/sinkhorn.py:149-208 (cost matrix :79-110, iterations :112-147).
//
// Two implementations of the same arithmetic:
//
//  * cluster path (sinkhorn_cluster_kernel): one 8-CTA thread-block cluster per descriptor pair.
//    Each CTA computes a 64-row slab of the similarity GEMM with FP32 FFMA straight into its
//    shared memory, so the (N+1)x(M+1) score matrix lives only in the cluster's distributed shared
//    memory for all iterations.  Row updates are CTA-local; column updates exchange one partial
//    sum per column per CTA through DSMEM (st to the peers' shared memory + one cluster barrier
//    per iteration).  Only P is written to HBM.
//    The duals are kept in the log domain (base 2), and every logsumexp is shifted by the dual
//    of the previous half-step instead of by the running maximum:
//        LSE_j(S_ij + v_j) = -u_i + log sum_j exp(S_ij + u_i + v_j)
//    After the first half-step every such exponent is the log of an entry of a row- or
//    column-normalised plan, hence bounded by log(N+M): no overflow, and the sums are bounded
//    below by 1/(N+M): no underflow.  The same exponentials, rescaled by mu_i/rowsum_i, are the
//    terms of the following column sum, so one exp per matrix entry per iteration suffices.
//    Rows whose sum leaves [1e-30, 1e30] (first half-step with a huge unused_score/epsilon) are
//    redone with the classic max shift.
//    Limits: squared-L2 cost, N <= 512, M <= 512, D % 16 == 0.
//
//  * generic path (separate kernels, score matrix in the P buffer in global memory, classic
//    max-shifted logsumexp): any N, M, D, L1 or L2 cost.  Used beyond the limits above
//    (e.g. K = 2048) and as an in-library cross-check.
#include <cooperative_groups.h>
#include <math_constants.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace om {

namespace {

// ------------------------------------------------------------------------------------------
// generic path
// ------------------------------------------------------------------------------------------
// squared norms of the rows; ovf (nullable): set when some |x| >= 60000 (out of range for the fp16 terms of cost_tc_kernel)
__global__ void __launch_bounds__(256) sqnorm_kernel(const float* d, long long rows, int D, float* out, unsigned int* ovf) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* p = d + row * D;
    float acc = 0.0f, mx = 0.0f;
    for (int k = threadIdx.x & 31; k < D; k += 32) {
        acc = fmaf(p[k], p[k], acc);
        mx = fmaxf(mx, fabsf(p[k]));
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) out[row] = acc;
    if (ovf != nullptr && !(mx < 60000.0f)) atomicOr(ovf, 1u);
}

// S[b][i][j] = -clamp(n1_i + n2_j - 2 a_i.b_j, 0) / eps for i<N, j<M; dustbin elsewhere (sinkhorn.py:98-103, :178-187)
// FP32 FFMA GEMM for matrices beyond the cluster kernel (K > 512): 128 x 128 output tile per CTA, 8 x 8 per thread
// (two 4-wide groups per side so that every operand read is one 16-byte shared-memory load), K in chunks of 16 with the
// next chunk prefetched into registers while the current one is multiplied.  Every output accumulates its D products in
// index order (one fmaf each), as the 64 x 64 / 4 x 4 kernel it replaces did: 26 -> ~50 TFLOP/s.
constexpr int CT = 128, CK = 16, CPITCH = CT + 4;
__global__ void __launch_bounds__(256, 2) cost_l2_kernel(const float* d1, const float* d2, const float* n1, const float* n2,
                                                      int N, int M, int D, float eps, float dustbin, int as_exp, float* S,
                                                      const unsigned int* only_if, int ld, size_t zs, int n1s, int n2s,
                                                      int flag_per_pair) {
    __shared__ __align__(16) float As[2][CK][CPITCH];
    // the tensor-core kernel did this launch's (flag_per_pair: this pair's) work unless the flag is set
    if (only_if != nullptr && only_if[flag_per_pair ? blockIdx.z : 0] == 0u) return;
    __shared__ __align__(16) float Bs[2][CK][CPITCH];
    const int z = blockIdx.z, i0 = blockIdx.y * CT, j0 = blockIdx.x * CT;
    const float* A = d1 + (size_t)z * N * D;
    const float* Bm = d2 + (size_t)z * M * D;
    const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
    const bool vec = (D & 3) == 0 && ((reinterpret_cast<uintptr_t>(d1) | reinterpret_cast<uintptr_t>(d2)) & 15) == 0;
    // loader: thread -> rows lr and lr + 64 of the tile, 4 consecutive k (lk ... lk+3) of the chunk
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    float4 ra[2], rb[2];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lr + 64 * h;
            const float* pa = A + (size_t)(i0 + r) * D + k0 + lk;
            const float* pb = Bm + (size_t)(j0 + r) * D + k0 + lk;
            const bool oka = i0 + r < N, okb = j0 + r < M;
            if (vec && k0 + lk + 3 < D) {
                ra[h] = oka ? __ldg(reinterpret_cast<const float4*>(pa)) : make_float4(0.f, 0.f, 0.f, 0.f);
                rb[h] = okb ? __ldg(reinterpret_cast<const float4*>(pb)) : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                float ta[4], tb[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    ta[q] = (oka && k0 + lk + q < D) ? __ldg(pa + q) : 0.0f;
                    tb[q] = (okb && k0 + lk + q < D) ? __ldg(pb + q) : 0.0f;
                }
                ra[h] = make_float4(ta[0], ta[1], ta[2], ta[3]);
                rb[h] = make_float4(tb[0], tb[1], tb[2], tb[3]);
            }
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lr + 64 * h;
            As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y; As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
            Bs[buf][lk + 0][r] = rb[h].x; Bs[buf][lk + 1][r] = rb[h].y; Bs[buf][lk + 2][r] = rb[h].z; Bs[buf][lk + 3][r] = rb[h].w;
        }
    };
    float acc[8][8] = {};
    const int nchunk = (D + CK - 1) / CK;
    fetch(0);
    stash(0);
    __syncthreads();
    for (int c = 0; c < nchunk; ++c) {
        const int buf = c & 1;
        if (c + 1 < nchunk) fetch((c + 1) * CK);
#pragma unroll
        for (int k = 0; k < CK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[r][q] = fmaf(av[r], bv[q], acc[r][q]);
        }
        if (c + 1 < nchunk) stash(buf ^ 1);
        __syncthreads();
    }
    float* Sz = S + (size_t)z * zs;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = i0 + (r >> 2) * 64 + ty * 4 + (r & 3);
        if (i > N) continue;
        const float n1i = i < N ? n1[(size_t)z * n1s + i] : 0.0f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int j = j0 + (q >> 2) * 64 + tx * 4 + (q & 3);
            if (j > M) continue;
            float v = dustbin;
            if (i < N && j < M) {
                const float cost = fmaxf(__fsub_rn(__fadd_rn(n1i, n2[(size_t)z * n2s + j]), __fmul_rn(2.0f, acc[r][q])), 0.0f);
                v = __fdiv_rn(-cost, eps);
            }
            Sz[(size_t)i * ld + j] = as_exp ? expf(v) : v;
        }
    }
}

__global__ void __launch_bounds__(256) cost_l1_kernel(const float* d1, const float* d2, int N, int M, int D, float eps,
                                                      float dustbin, int as_exp, float* S) {
    const int z = blockIdx.z;
    const int j = blockIdx.x * 32 + (threadIdx.x & 31), i = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (i > N || j > M) return;
    float v = dustbin;
    if (i < N && j < M) {
        const float* a = d1 + ((size_t)z * N + i) * D;
        const float* b = d2 + ((size_t)z * M + j) * D;
        float acc = 0.0f;
        for (int k = 0; k < D; ++k) acc += fabsf(a[k] - b[k]);            // sinkhorn.py:106-108
        v = __fdiv_rn(-acc, eps);
    }
    S[(size_t)z * (N + 1) * (M + 1) + (size_t)i * (M + 1) + j] = as_exp ? expf(v) : v;
}

// u_i = log_mu_i - logsumexp_j(S_ij + v_j)   (sinkhorn.py:140), one warp per row
__global__ void __launch_bounds__(256) row_pass_kernel(const float* S, const float* v, float* u, int N, int M,
                                                       float log_m) {
    const int z = blockIdx.y;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i > N) return;
    const float* row = S + (size_t)z * (N + 1) * (M + 1) + (size_t)i * (M + 1);
    const float* vz = v + (size_t)z * (M + 1);
    float m = -CUDART_INF_F;
    for (int j = lane; j <= M; j += 32) m = fmaxf(m, row[j] + vz[j]);
    m = warp_max(m);
    float s = 0.0f;
    for (int j = lane; j <= M; j += 32) s += expf((row[j] + vz[j]) - m);
    s = warp_sum(s);
    if (lane == 0) u[(size_t)z * (N + 1) + i] = (i == N ? log_m : 0.0f) - (logf(s) + m);
}

// v_j = log_nu_j - logsumexp_i(S_ij + u_i)   (sinkhorn.py:142); 32 columns x 8 row slices per CTA
__global__ void __launch_bounds__(256) col_pass_kernel(const float* S, const float* u, float* v, int N, int M,
                                                       float log_n) {
    __shared__ float sm[8][32], ss[8][32];
    const int z = blockIdx.y;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + tx;
    const float* Sz = S + (size_t)z * (N + 1) * (M + 1);
    const float* uz = u + (size_t)z * (N + 1);
    float m = -CUDART_INF_F, s = 0.0f;
    if (j <= M) {
        for (int i = ty; i <= N; i += 8) {
            const float t = Sz[(size_t)i * (M + 1) + j] + uz[i];
            if (t > m) { s = s * expf(m - t) + 1.0f; m = t; }
            else s += expf(t - m);
        }
    }
    sm[ty][tx] = m;
    ss[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && j <= M) {
        float mm = sm[0][tx];
        for (int k = 1; k < 8; ++k) mm = fmaxf(mm, sm[k][tx]);
        float tot = 0.0f;
        for (int k = 0; k < 8; ++k)
            if (ss[k][tx] > 0.0f) tot += ss[k][tx] * expf(sm[k][tx] - mm);
        v[(size_t)z * (M + 1) + j] = (j == M ? log_n : 0.0f) - (logf(tot) + mm);
    }
}

// P = exp(S + u + v) in place (sinkhorn.py:145, :206)
__global__ void __launch_bounds__(256) finalize_kernel(float* S, const float* u, const float* v, int N, int M) {
    const int z = blockIdx.y;
    const size_t n = (size_t)(N + 1) * (M + 1);
    float* Sz = S + (size_t)z * n;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (size_t)gridDim.x * 256) {
        const int i = (int)(e / (M + 1)), j = (int)(e % (M + 1));
        Sz[e] = expf((Sz[e] + u[(size_t)z * (N + 1) + i]) + v[(size_t)z * (M + 1) + j]);
    }
}

// ---- scaling form of the same iteration for matrices that do not fit a cluster (K = exp(S) in global memory / L2) ----
// a_i = mu_i / sum_j K_ij b_j,  b_j = nu_j / sum_i K_ij a_i  (== sinkhorn.py:138-142 with a = exp(u), b = exp(v)),
// used while exp(-unused/eps) is far from underflow.  One sweep of K per iteration: a CTA owns XR rows, computes their
// row sums (one warp per row) and at once their contribution to every column sum (the rows are re-read from L1);
// the per-CTA column partials are combined in a fixed order by xd_col_kernel (no floating-point atomics).
constexpr int XR = 16;

__global__ void __launch_bounds__(XR * 32) xd_row_kernel(const float* K, const float* b, float* a, float* Tpart, int N, int M,
                                                         float mu_dust) {
    __shared__ float sA[XR];
    const int z = blockIdx.y, blk = blockIdx.x;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i0 = blk * XR, i = i0 + w;
    const float* Kz = K + (size_t)z * (N + 1) * (M + 1);
    const float* bz = b + (size_t)z * (M + 1);
    float av = 0.0f;
    if (i <= N) {
        const float* row = Kz + (size_t)i * (M + 1);
        float s0 = 0.0f, s1 = 0.0f;
        int j = lane;
        for (; j + 32 <= M; j += 64) { s0 = fmaf(row[j], bz[j], s0); s1 = fmaf(row[j + 32], bz[j + 32], s1); }
        for (; j <= M; j += 32) s0 = fmaf(row[j], bz[j], s0);
        const float rs = warp_sum(s0 + s1);
        av = __fdividef(i == N ? mu_dust : 1.0f, rs);
        if (lane == 0) a[(size_t)z * (N + 1) + i] = av;
    }
    if (lane == 0) sA[w] = av;
    __syncthreads();
    const int rows = min(XR, N + 1 - i0);
    float* tp = Tpart + ((size_t)z * gridDim.x + blk) * (M + 1);
    for (int j = threadIdx.x; j <= M; j += XR * 32) {
        float t = 0.0f;
        for (int r = 0; r < rows; ++r) t = fmaf(Kz[(size_t)(i0 + r) * (M + 1) + j], sA[r], t);
        tp[j] = t;
    }
}

// 32 columns per CTA, 8 warps: warp g adds the partials of row blocks g, g + 8, ... of its 32 columns (coalesced 128-byte rows),
// the eight group sums are combined in a fixed order.  (One thread per column over all row blocks left 72 CTAs of serial
// 129-term sums at K = 2048, batch 8: 16.5 us per iteration.)
__global__ void __launch_bounds__(256) xd_col_kernel(const float* Tpart, float* b, int nblk, int M, float nu_dust) {
    __shared__ float sT[8][33];
    const int z = blockIdx.y;
    const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + c;
    float t0 = 0.0f, t1 = 0.0f;
    if (j <= M) {
        const float* tp = Tpart + (size_t)z * nblk * (M + 1) + j;
        int q = g;
        for (; q + 8 < nblk; q += 16) { t0 += tp[(size_t)q * (M + 1)]; t1 += tp[(size_t)(q + 8) * (M + 1)]; }
        if (q < nblk) t0 += tp[(size_t)q * (M + 1)];
    }
    sT[g][c] = t0 + t1;
    __syncthreads();
    if (g == 0 && j <= M) {
        float t = sT[0][c];
#pragma unroll
        for (int k = 1; k < 8; ++k) t += sT[k][c];
        b[(size_t)z * (M + 1) + j] = __fdividef(j == M ? nu_dust : 1.0f, t);
    }
}

// P = a_i K_ij b_j in place (sinkhorn.py:145, :206)
__global__ void __launch_bounds__(256) xd_finalize_kernel(float* K, const float* a, const float* b, int N, int M) {
    const int z = blockIdx.y;
    const size_t n = (size_t)(N + 1) * (M + 1);
    float* Kz = K + (size_t)z * n;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (size_t)gridDim.x * 256) {
        const int i = (int)(e / (M + 1)), j = (int)(e % (M + 1));
        Kz[e] = a[(size_t)z * (N + 1) + i] * Kz[e] * b[(size_t)z * (M + 1) + j];
    }
}

__global__ void __launch_bounds__(256) fill_kernel(float* p, size_t n, float v) {
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (size_t)gridDim.x * 256) p[e] = v;
}

struct SkWs {
    float *n1, *n2, *u, *v, *tpart;
};

SkWs carve_sk(void* ws, int B, int N, int M) {
    SkWs w;
    char* p = (char*)ws;
    w.n1 = (float*)p; p += align_up((size_t)B * N * sizeof(float));
    w.n2 = (float*)p; p += align_up((size_t)B * M * sizeof(float));
    w.u = (float*)p;  p += align_up((size_t)B * (N + 1) * sizeof(float));
    w.v = (float*)p;  p += align_up((size_t)B * (M + 1) * sizeof(float));
    w.tpart = (float*)p;
    return w;
}

int g_generic_allow_scaling = 1;   // test hook: 0 forces the log-domain generic kernels
int g_generic_tc = 1;              // test hook: 0 forces the FP32 FFMA cost kernel on the generic path

int sinkhorn_generic(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float eps,
                     float dustbin, int l1, float* P, void* ws, cudaStream_t st) {
    // (Measured twice, rounds 1 and 2: running cost -> iterations -> P in L2-sized groups of pairs, so that the sweeps of
    // K = 2048 matrices come from L2 instead of HBM, makes this stage 3 % faster timed alone and the whole step 3 % slower.)
    const SkWs w = carve_sk(ws, B, N, M);
    // scaling form while exp(dustbin) = exp(-unused/eps) stays far from underflow (2^-60), as in the cluster kernel
    const int scaling = g_generic_allow_scaling && dustbin <= 0.0f && -(double)dustbin * 1.4426950408889634 <= 60.0;
    if (l1) {
        cost_l1_kernel<<<dim3((M + 32) / 32, (N + 8) / 8, B), 256, 0, st>>>(d1, d2, N, M, D, eps, dustbin, scaling, P);
        OM_AFTER_LAUNCH();
    } else {
        // similarity GEMM on tcgen05 (two-term fp16 split) when D allows; the FP32 FFMA kernel runs behind it and only
        // does work when some |x| >= 60000 was seen (flag in the first word of w.u, which the iterations overwrite later)
        const bool tc = g_generic_tc && D % 32 == 0;
        unsigned int* ovf = tc ? reinterpret_cast<unsigned int*>(w.u) : nullptr;
        if (tc) OM_CUDA(cudaMemsetAsync(ovf, 0, sizeof(unsigned int), st));
        sqnorm_kernel<<<(unsigned)(((long long)B * N + 7) / 8), 256, 0, st>>>(d1, (long long)B * N, D, w.n1, ovf);
        OM_AFTER_LAUNCH();
        sqnorm_kernel<<<(unsigned)(((long long)B * M + 7) / 8), 256, 0, st>>>(d2, (long long)B * M, D, w.n2, ovf);
        OM_AFTER_LAUNCH();
        if (tc) OM_TRY(cost_tc_launch(d1, d2, w.n1, w.n2, B, N, M, D, eps, dustbin, scaling, P, ovf, st));
        cost_l2_kernel<<<dim3((M + CT) / CT, (N + CT) / CT, B), 256, 0, st>>>(d1, d2, w.n1, w.n2, N, M, D, eps, dustbin, scaling, P, ovf,
                                                                               M + 1, (size_t)(N + 1) * (M + 1), N, M, 0);
        OM_AFTER_LAUNCH();
    }
    if (scaling) {
        const int nblk = (N + 1 + XR - 1) / XR;
        fill_kernel<<<64, 256, 0, st>>>(w.v, (size_t)B * (M + 1), 1.0f);              // b = exp(v) = 1
        OM_AFTER_LAUNCH();
        for (int it = 0; it < iterations; ++it) {
            xd_row_kernel<<<dim3(nblk, B), XR * 32, 0, st>>>(P, w.v, w.u, w.tpart, N, M, (float)M);   // mu_N = M (:197-198)
            OM_AFTER_LAUNCH();
            xd_col_kernel<<<dim3((M + 32) / 32, B), 256, 0, st>>>(w.tpart, w.v, nblk, M, (float)N);  // nu_M = N (:199-200)
            OM_AFTER_LAUNCH();
        }
        const size_t n = (size_t)(N + 1) * (M + 1);
        xd_finalize_kernel<<<dim3((unsigned)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024), B), 256, 0, st>>>(P, w.u, w.v, N, M);
        OM_AFTER_LAUNCH();
        return OM_OK;
    }
    OM_CUDA(cudaMemsetAsync(w.v, 0, (size_t)B * (M + 1) * sizeof(float), st));
    const float log_m = (float)log((double)M), log_n = (float)log((double)N);       // sinkhorn.py:197-198
    for (int it = 0; it < iterations; ++it) {
        row_pass_kernel<<<dim3((N + 8) / 8, B), 256, 0, st>>>(P, w.v, w.u, N, M, log_m);
        OM_AFTER_LAUNCH();
        col_pass_kernel<<<dim3((M + 32) / 32, B), 256, 0, st>>>(P, w.u, w.v, N, M, log_n);
        OM_AFTER_LAUNCH();
    }
    const size_t n = (size_t)(N + 1) * (M + 1);
    finalize_kernel<<<dim3((unsigned)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024), B), 256, 0, st>>>(P, w.u, w.v, N, M);
    OM_AFTER_LAUNCH();
    return OM_OK;
}

// ------------------------------------------------------------------------------------------
// cluster path
// ------------------------------------------------------------------------------------------
constexpr int CL = 8;              // CTAs per cluster == per descriptor pair
constexpr int CNT = 256;           // threads per CTA
constexpr int RPC = 64;            // real rows per CTA
constexpr int MAXM = 512;
constexpr int NCOL = 544;          // 17 columns per lane
constexpr int SPITCH = NCOL;       // floats per score row: M+1 used, the rest is -inf padding (exp -> 0)
constexpr int CPL = NCOL / 32;     // 17
constexpr int KC = 16;             // GEMM k chunk
constexpr int AP = 68;             // pitch of the A chunk [KC][64]
constexpr int BP = 516;            // pitch of the B chunk [KC][512]

// shared-memory plan (floats)
constexpr int OFF_S = 0;                                   // (RPC+1) x SPITCH
constexpr int OFF_U = OFF_S + (RPC + 1) * SPITCH;          // u2 per local row (+pad)
constexpr int OFF_V = OFF_U + 72;                          // v2 per column
constexpr int OFF_RECV = OFF_V + NCOL;                     // [2][CL][NCOL] partials received from the cluster
constexpr int OFF_X = OFF_RECV + 2 * CL * NCOL;            // union: GEMM chunks | per-warp column accumulators
constexpr int X_GEMM = KC * AP + KC * BP + 64 + MAXM;      // A chunk, B chunk, n1, n2
constexpr int X_ITER = (CNT / 32) * NCOL;
constexpr int X_SIZE = X_GEMM > X_ITER ? X_GEMM : X_ITER;
constexpr int CLUSTER_SMEM_FLOATS = OFF_X + X_SIZE;

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

struct ClusterArgs {
    const float* d1;
    const float* d2;
    int N, M, D;
    int iterations;
    float scale2;       // log2(e)/eps: cost -> base-2 log score
    float dustbin2;     // (-unused/eps) * log2(e)
    float* P;
};

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(CNT, 1) sinkhorn_cluster_kernel(ClusterArgs a) {
    extern __shared__ __align__(16) float sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int z = blockIdx.x / CL;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = a.N, M = a.M, D = a.D;

    float* sS = sm + OFF_S;
    float* sU = sm + OFF_U;
    float* sV = sm + OFF_V;
    float* sRecv = sm + OFF_RECV;
    float* sX = sm + OFF_X;

    const int r0 = rank * RPC;                                  // first real row of this CTA
    const int nreal = max(0, min(RPC, N - r0));                 // real rows held here
    const bool has_dust = rank == CL - 1;                       // dustbin row lives in the last CTA
    const int nloc = nreal + (has_dust ? 1 : 0);

    // ---------------- similarity GEMM: sS[li][j] = -max(n1+n2-2 a.b, 0) * scale2 -------------------
    {
        float* sA = sX;                    // [KC][AP]
        float* sB = sA + KC * AP;          // [KC][BP]
        float* sN1 = sB + KC * BP;         // [64]
        float* sN2 = sN1 + 64;             // [512]
        const float* A = a.d1 + ((size_t)z * N + r0) * D;
        const float* Bm = a.d2 + (size_t)z * M * D;
        const int tr = warp, tc = lane;    // 8 rows x 16 cols per thread
        float acc[8][16];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[r][c] = 0.0f;
        const int lrow = tid >> 2, lkq = tid & 3;               // loader coordinates
        float na = 0.0f, nb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k0 = 0; k0 < D; k0 += KC) {
            {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lrow < nreal) v = __ldg(reinterpret_cast<const float4*>(A + (size_t)lrow * D + k0 + 4 * lkq));
                na = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, na))));
                sA[(4 * lkq + 0) * AP + lrow] = v.x; sA[(4 * lkq + 1) * AP + lrow] = v.y;
                sA[(4 * lkq + 2) * AP + lrow] = v.z; sA[(4 * lkq + 3) * AP + lrow] = v.w;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int j = lrow + 64 * q;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < M) v = __ldg(reinterpret_cast<const float4*>(Bm + (size_t)j * D + k0 + 4 * lkq));
                nb[q] = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, nb[q]))));
                sB[(4 * lkq + 0) * BP + j] = v.x; sB[(4 * lkq + 1) * BP + j] = v.y;
                sB[(4 * lkq + 2) * BP + j] = v.z; sB[(4 * lkq + 3) * BP + j] = v.w;
            }
            __syncthreads();
#pragma unroll 4
            for (int k = 0; k < KC; ++k) {
                const float4 a0 = *reinterpret_cast<const float4*>(sA + k * AP + tr * 8);
                const float4 a1 = *reinterpret_cast<const float4*>(sA + k * AP + tr * 8 + 4);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 b = *reinterpret_cast<const float4*>(sB + k * BP + 128 * q + 4 * tc);
                    const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int r = 0; r < 8; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[r][4 * q + c] = fmaf(av[r], bv[c], acc[r][4 * q + c]);
                }
            }
            __syncthreads();
        }
        // squared norms: 4 loader threads share a row (sinkhorn.py:98-99)
        na += __shfl_xor_sync(0xffffffffu, na, 1);
        na += __shfl_xor_sync(0xffffffffu, na, 2);
        if (lkq == 0) sN1[lrow] = na;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float v = nb[q];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (lkq == 0) sN2[lrow + 64 * q] = v;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int li = tr * 8 + r;
            if (li >= nreal) continue;
            const float n1 = sN1[li];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = 128 * q + 4 * tc;
                float o[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float cost = fmaxf(__fsub_rn(__fadd_rn(n1, sN2[j + c]), __fmul_rn(2.0f, acc[r][4 * q + c])), 0.0f);
                    o[c] = __fmul_rn(-cost, a.scale2);
                }
                *reinterpret_cast<float4*>(sS + li * SPITCH + j) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
        __syncthreads();
        // dustbin column / row (sinkhorn.py:182-187); columns beyond M are -inf so that they drop
        // out of every sum without a predicate
        for (int e2 = tid; e2 < nloc * (NCOL - M); e2 += CNT) {
            const int li = e2 / (NCOL - M), j = M + e2 % (NCOL - M);
            sS[li * SPITCH + j] = (j == M) ? a.dustbin2 : -CUDART_INF_F;
        }
        if (has_dust)
            for (int j = tid; j < M; j += CNT) sS[nreal * SPITCH + j] = a.dustbin2;
    }
    for (int i = tid; i < 72; i += CNT) sU[i] = 0.0f;
    for (int i = tid; i < NCOL; i += CNT) sV[i] = 0.0f;
    __syncthreads();

    const float log2_m = log2f((float)M), log2_n = log2f((float)N);
    float* sCW = sX;                                              // [8 warps][NCOL], aliases the GEMM chunks
    float vreg[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) vreg[k] = 0.0f;

    // where this rank's partial sums land in each peer's receive buffer (parity 0)
    float* peer_slot[CL];
#pragma unroll
    for (int dst = 0; dst < CL; ++dst) peer_slot[dst] = cluster.map_shared_rank(sRecv + rank * NCOL, dst);

    // make sure every CTA of the cluster is running before the first DSMEM store
    cluster.sync();

    for (int it = 0; it < a.iterations; ++it) {
        float colacc[CPL];
#pragma unroll
        for (int k = 0; k < CPL; ++k) colacc[k] = 0.0f;
        for (int li = warp; li < nloc; li += CNT / 32) {
            const bool dust_row = has_dust && li == nreal;
            const float mu = dust_row ? (float)M : 1.0f;
            const float u_old = sU[li];
            const float* row = sS + li * SPITCH;
            float e[CPL];
            float rs = 0.0f;
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                e[k] = ex2((row[lane + 32 * k] + vreg[k]) + u_old);
                rs += e[k];
            }
            rs = warp_sum(rs);
            float u_new, f;
            if (rs >= 1e-30f && rs <= 1e30f) {
                f = __fdividef(mu, rs);
                u_new = u_old + ((dust_row ? log2_m : 0.0f) - lg2(rs));
            } else {
                // classic max-shifted logsumexp for this row (sinkhorn.py:140)
                float m = -CUDART_INF_F;
#pragma unroll
                for (int k = 0; k < CPL; ++k) m = fmaxf(m, row[lane + 32 * k] + vreg[k]);
                m = warp_max(m);
                rs = 0.0f;
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    e[k] = ex2((row[lane + 32 * k] + vreg[k]) - m);
                    rs += e[k];
                }
                rs = warp_sum(rs);
                f = __fdividef(mu, rs);
                u_new = (dust_row ? log2_m : 0.0f) - (m + lg2(rs));
            }
#pragma unroll
            for (int k = 0; k < CPL; ++k) colacc[k] = fmaf(e[k], f, colacc[k]);   // exp2(S + u_new + v)
            if (lane == 0) sU[li] = u_new;
        }
#pragma unroll
        for (int k = 0; k < CPL; ++k) sCW[warp * NCOL + lane + 32 * k] = colacc[k];
        __syncthreads();
        // CTA partial per column, pushed as float4 into every cluster member's receive slot for this
        // rank (DSMEM stores; the peers' base addresses were mapped once before the loop)
        if (tid < NCOL / 4) {
            float4 s4 = *reinterpret_cast<const float4*>(sCW + 4 * tid);
#pragma unroll
            for (int w = 1; w < CNT / 32; ++w) {
                const float4 t4 = *reinterpret_cast<const float4*>(sCW + w * NCOL + 4 * tid);
                s4.x += t4.x; s4.y += t4.y; s4.z += t4.z; s4.w += t4.w;
            }
            const int off = (it & 1) * CL * NCOL + 4 * tid;
#pragma unroll
            for (int dst = 0; dst < CL; ++dst) *reinterpret_cast<float4*>(peer_slot[dst] + off) = s4;
        }
        cluster.sync();
        // v_j += log_nu_j - log(colsum_j)   (sinkhorn.py:142 with the shift -v_j)
        const float* recv = sRecv + (it & 1) * CL * NCOL;
        for (int c = tid; c <= M; c += CNT) {
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
    for (int li = warp; li < nloc; li += CNT / 32) {
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
    // no CTA may exit while a peer can still write into its shared memory
    cluster.sync();
}

int sinkhorn_cluster(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float eps,
                     float unused, float* P, cudaStream_t st) {
    ClusterArgs a{};
    a.d1 = d1; a.d2 = d2; a.N = N; a.M = M; a.D = D; a.iterations = iterations; a.P = P;
    const double log2e = 1.4426950408889634;
    a.scale2 = (float)(log2e / (double)eps);
    a.dustbin2 = (float)((-(double)unused / (double)eps) * log2e);
    const size_t smem = (size_t)CLUSTER_SMEM_FLOATS * sizeof(float);
    OM_TRY(set_smem(sinkhorn_cluster_kernel, smem));
    sinkhorn_cluster_kernel<<<B * CL, CNT, smem, st>>>(a);
    OM_AFTER_LAUNCH();
    return OM_OK;
}

// 0: tcgen05 cluster kernel (scaling-form loop when safe), 1: FFMA cluster kernel, 2: generic kernels,
// 3: tcgen05 cluster kernel with the log-domain loop forced, 4: tcgen05 cluster kernel with the 3xTF32 GEMM forced,
// 5: generic kernels with the log-domain loop forced (2 = generic kernels, scaling form when safe),
// 7: generic kernels with the FP32 FFMA cost GEMM (2 / 5: cost GEMM on tcgen05 when D % 32 == 0)
// 0 takes the hybrid-resident cluster kernel (sinkhorn_hy.cu) whenever it is eligible; 8: the 8-CTA tcgen05 kernel as 0 did
// before the hybrid kernel existed (K <= 512) / the generic kernels (K > 512)
// 9: the streaming kernels of sinkhorn_xl.cu at any size they are eligible for (0 takes them beyond the hybrid kernel's sizes)
int g_sinkhorn_variant = 0;

}  // namespace

bool sinkhorn_xl_eligible(int N, int M, int D, float eps, float unused, int distance_l1);
size_t sinkhorn_xl_workspace_bytes(int B, int N, int M, int D);
int sinkhorn_xl_launch(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float eps, float unused,
                       float* P, void* ws, size_t ws_bytes, cudaStream_t st);
extern int g_hy_allow_16;

int cost_l2_fallback_launch(const float* d1, const float* d2, const float* n1, int n1_stride, const float* n2, int n2_stride, int B,
                            int N, int M, int D, float eps, float dustbin, float* S, int ld, size_t zstride,
                            const unsigned int* per_pair_flag, cudaStream_t st) {
    cost_l2_kernel<<<dim3((M + CT) / CT, (N + CT) / CT, B), 256, 0, st>>>(d1, d2, n1, n2, N, M, D, eps, dustbin, 1, S, per_pair_flag, ld,
                                                                           zstride, n1_stride, n2_stride, 1);
    OM_AFTER_LAUNCH();
    return OM_OK;
}

// does this problem size ask for the streaming kernels' workspace (beyond the hybrid kernel's sizes, or forced by the test hook)?
static bool xl_sized(int N, int M, int D) {
    if (D % 32 != 0) return false;
    return g_sinkhorn_variant == 9 || N > 1024 || M > 1024 || (!g_hy_allow_16 && (N > 512 || M > 512));
}

size_t sinkhorn_workspace_bytes(int B, int N, int M, int D) {
    (void)D;
    if (B <= 0 || N <= 0 || M <= 0) return 0;
    const size_t generic = align_up((size_t)B * N * sizeof(float)) + align_up((size_t)B * M * sizeof(float)) +
           align_up((size_t)B * (N + 1) * sizeof(float)) + align_up((size_t)B * (M + 1) * sizeof(float)) +
           align_up((size_t)B * ((N + 1 + XR - 1) / XR) * (M + 1) * sizeof(float));   // column partials of the scaling-form generic path
    const size_t hy = D % 32 == 0 ? sinkhorn_hy_workspace_bytes(B, N, M, D) : 0;       // packed fp16 operands of the hybrid kernel
    const size_t xl = xl_sized(N, M, D) ? sinkhorn_xl_workspace_bytes(B, N, M, D) : 0; // packed operands + aligned K of the streaming kernels
    const size_t m = generic > hy ? generic : hy;
    return m > xl ? m : xl;
}

static size_t generic_workspace_bytes(int B, int N, int M) {
    return align_up((size_t)B * N * sizeof(float)) + align_up((size_t)B * M * sizeof(float)) +
           align_up((size_t)B * (N + 1) * sizeof(float)) + align_up((size_t)B * (M + 1) * sizeof(float)) +
           align_up((size_t)B * ((N + 1 + XR - 1) / XR) * (M + 1) * sizeof(float));
}

bool sinkhorn_routes_to_hy(int B, int N, int M, int D, float epsilon, float unused_score, int distance_l1, const void* ws,
                           size_t ws_bytes) {
    return g_sinkhorn_variant == 0 && B > 0 && B <= 65535 && (long long)B * 16 < (1ll << 31) &&
           sinkhorn_hy_eligible(N, M, D, epsilon, unused_score, distance_l1) && ws != nullptr &&
           ws_bytes >= sinkhorn_hy_workspace_bytes(B, N, M, D);
}

// will sinkhorn_launch take the streaming kernels (sinkhorn_xl.cu) for this problem?  (the fused matcher then packs each image's
// descriptors on that image's stream)
bool sinkhorn_routes_to_xl(int B, int N, int M, int D, float epsilon, float unused_score, int distance_l1, const void* ws,
                           size_t ws_bytes) {
    if (B <= 0 || B > 65535 || N <= 0 || M <= 0 || D <= 0 || !(epsilon > 0.0f)) return false;
    if (sinkhorn_routes_to_hy(B, N, M, D, epsilon, unused_score, distance_l1, ws, ws_bytes)) return false;
    const bool fast = !distance_l1 && N <= RPC * CL && M <= MAXM && D % KC == 0 && g_sinkhorn_variant != 2 &&
                      g_sinkhorn_variant != 5 && g_sinkhorn_variant != 7 && g_sinkhorn_variant != 9 && (long long)B * CL < (1ll << 31);
    if (fast) return false;
    return (g_sinkhorn_variant == 0 || g_sinkhorn_variant == 9) && sinkhorn_xl_eligible(N, M, D, epsilon, unused_score, distance_l1) &&
           ws != nullptr && ws_bytes >= sinkhorn_xl_workspace_bytes(B, N, M, D);
}

int sinkhorn_launch(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float epsilon,
                    float unused_score, int distance_l1, float* P, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (d1 == nullptr || d2 == nullptr || P == nullptr) return OM_ERR_NULL;
    if (B <= 0 || N <= 0 || M <= 0 || D <= 0) return OM_ERR_SHAPE;
    if (iterations <= 0 || !(epsilon > 0.0f)) return OM_ERR_PARAM;        // sinkhorn.py:66-69
    if (B > 65535) return OM_ERR_LIMIT;
    const bool fast = !distance_l1 && N <= RPC * CL && M <= MAXM && D % KC == 0 && g_sinkhorn_variant != 2 &&
                      g_sinkhorn_variant != 5 && g_sinkhorn_variant != 7 && g_sinkhorn_variant != 9 && (long long)B * CL < (1ll << 31);
    g_generic_allow_scaling = g_sinkhorn_variant != 5;
    g_generic_tc = g_sinkhorn_variant != 7;
    if (g_sinkhorn_variant == 0 && (long long)B * 16 < (1ll << 31) &&
        sinkhorn_hy_eligible(N, M, D, epsilon, unused_score, distance_l1) && ws != nullptr &&
        ws_bytes >= sinkhorn_hy_workspace_bytes(B, N, M, D))
        return sinkhorn_hy_launch(d1, d2, B, N, M, D, iterations, epsilon, unused_score, P, nullptr, ws, ws_bytes, st);
    if (fast && (g_sinkhorn_variant == 0 || g_sinkhorn_variant == 8 || g_sinkhorn_variant == 3 || g_sinkhorn_variant == 4)) {
        g_tc_allow_scaling = g_sinkhorn_variant != 3;
        g_tc_allow_f16 = g_sinkhorn_variant != 4;
        return sinkhorn_cluster_tc(d1, d2, B, N, M, D, iterations, epsilon, unused_score, P, st);
    }
    if (fast) return sinkhorn_cluster(d1, d2, B, N, M, D, iterations, epsilon, unused_score, P, st);
    if ((g_sinkhorn_variant == 0 || g_sinkhorn_variant == 9) && sinkhorn_xl_eligible(N, M, D, epsilon, unused_score, distance_l1) &&
        ws != nullptr && ws_bytes >= sinkhorn_xl_workspace_bytes(B, N, M, D))
        return sinkhorn_xl_launch(d1, d2, B, N, M, D, iterations, epsilon, unused_score, P, ws, ws_bytes, st);
    if (ws == nullptr || ws_bytes < generic_workspace_bytes(B, N, M)) return OM_ERR_WORKSPACE;
    // dustbin score is computed in double by the reference (python floats) and cast once, sinkhorn.py:182
    const float dustbin = (float)(-(double)unused_score / (double)epsilon);
    return sinkhorn_generic(d1, d2, B, N, M, D, iterations, epsilon, dustbin, distance_l1, P, ws, st);
}

size_t sinkhorn_ex_workspace_bytes(int B, int N, int M, int D) {
    if (B <= 0 || N <= 0 || M <= 0) return 0;
    return align_up(sinkhorn_workspace_bytes(B, N, M, D)) + align_up((size_t)B * (N + 1) * (M + 1) * sizeof(float)) +
           align_up(mutual_matches_workspace_bytes(B, N, M));
}

int sinkhorn_ex_launch(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float epsilon,
                       float unused_score, int distance_l1, float* P, const SinkhornEpilogue& e, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
    if (!e.any()) {
        if (P == nullptr) return OM_ERR_NULL;
        return sinkhorn_launch(d1, d2, B, N, M, D, iterations, epsilon, unused_score, distance_l1, P, ws, ws_bytes, st);
    }
    if (d1 == nullptr || d2 == nullptr) return OM_ERR_NULL;
    if (B <= 0 || N <= 0 || M <= 0 || D <= 0) return OM_ERR_SHAPE;
    if (iterations <= 0 || !(epsilon > 0.0f)) return OM_ERR_PARAM;
    if (B > 65535) return OM_ERR_LIMIT;
    if (e.matches && (e.kpts1 == nullptr || e.kpts2 == nullptr || e.mk1 == nullptr || e.mk2 == nullptr || e.mscores == nullptr ||
                      e.mvalid == nullptr))
        return OM_ERR_NULL;
    if (e.matches && e.max_matches <= 0) return OM_ERR_SHAPE;
    if (e.filters && e.filter_valid == nullptr) return OM_ERR_NULL;
    // fused: the epilogue runs on the P values in the cluster kernel's registers; P is written only if the caller wants it
    if (g_sinkhorn_variant == 0 || g_sinkhorn_variant == 8) g_tc_allow_scaling = g_tc_allow_f16 = 1;     // (test variants leave these switched)
    if (g_sinkhorn_variant == 0 && (long long)B * 16 < (1ll << 31) &&
        sinkhorn_hy_eligible(N, M, D, epsilon, unused_score, distance_l1) && ws != nullptr &&
        ws_bytes >= sinkhorn_hy_workspace_bytes(B, N, M, D))
        return sinkhorn_hy_launch(d1, d2, B, N, M, D, iterations, epsilon, unused_score, P, &e, ws, ws_bytes, st);
    if ((g_sinkhorn_variant == 0 || g_sinkhorn_variant == 8) && (long long)B * CL < (1ll << 31) &&
        sinkhorn_epilogue_can_fuse(N, M, D, epsilon, unused_score, distance_l1))
        return sinkhorn_cluster_tc_epi(d1, d2, B, N, M, D, iterations, epsilon, unused_score, P, e, st);
    // separate kernels on a stored P (more than 512 keypoints, L1 cost, log-domain loop, test variants)
    if (ws == nullptr || ws_bytes < sinkhorn_ex_workspace_bytes(B, N, M, D)) return OM_ERR_WORKSPACE;
    char* c = (char*)ws;
    void* sink_ws = c;
    c += align_up(sinkhorn_workspace_bytes(B, N, M, D));
    float* Pbuf = P != nullptr ? P : (float*)c;
    c += align_up((size_t)B * (N + 1) * (M + 1) * sizeof(float));
    OM_TRY(sinkhorn_launch(d1, d2, B, N, M, D, iterations, epsilon, unused_score, distance_l1, Pbuf, sink_ws,
                           sinkhorn_workspace_bytes(B, N, M, D), st));
    if (e.filters) OM_TRY(filter_rows_launch(Pbuf, B, N, M, e.ratio_threshold, e.dustbin_margin, e.filter_valid, st));
    if (e.scores0 != nullptr || e.scores1 != nullptr) OM_TRY(sinkhorn_scores_launch(Pbuf, B, N, M, e.scores0, e.scores1, st));
    if (e.matches)
        OM_TRY(mutual_matches_launch(Pbuf, e.kpts1, e.kpts2, B, N, M, e.max_matches, e.match_threshold, e.mk1, e.mk2, e.mscores,
                                     e.mvalid, c, mutual_matches_workspace_bytes(B, N, M), st));
    return OM_OK;
}

SinkhornEpilogue epilogue_from_outputs(const om_sinkhorn_outputs* o) {
    SinkhornEpilogue e;
    if (o == nullptr) return e;
    e.scores0 = o->scores0; e.scores1 = o->scores1;
    e.filters = o->filters; e.ratio_threshold = o->ratio_threshold; e.dustbin_margin = o->dustbin_margin;
    e.filter_valid = o->filter_valid;
    e.matches = o->matches; e.kpts1 = o->kpts1; e.kpts2 = o->kpts2; e.max_matches = o->max_matches;
    e.match_threshold = o->match_threshold; e.mk1 = o->matched_kpts1; e.mk2 = o->matched_kpts2; e.mscores = o->match_scores;
    e.mvalid = o->match_valid;
    return e;
}

}  // namespace om

using namespace om;

extern "C" size_t om_sinkhorn_ex_workspace_bytes(int B, int N, int M, int D) { return sinkhorn_ex_workspace_bytes(B, N, M, D); }

extern "C" int om_sinkhorn_ex_f32(const float* desc1, const float* desc2, int B, int N, int M, int D, int iterations,
                                  float epsilon, float unused_score, int distance_l1, const om_sinkhorn_outputs* out, void* ws,
                                  size_t ws_bytes, void* stream) {
    OM_ON_DEVICE_OF(desc1);
    if (out == nullptr) return OM_ERR_NULL;
    return sinkhorn_ex_launch(desc1, desc2, B, N, M, D, iterations, epsilon, unused_score, distance_l1, out->probs,
                              epilogue_from_outputs(out), ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" void om_debug_force_generic_sinkhorn(int on) { g_sinkhorn_variant = on ? 2 : 0; }
extern "C" void om_debug_sinkhorn_variant(int variant) { g_sinkhorn_variant = variant; }

extern "C" size_t om_sinkhorn_workspace_bytes(int B, int N, int M, int D) { return sinkhorn_workspace_bytes(B, N, M, D); }

extern "C" int om_sinkhorn_f32(const float* desc1, const float* desc2, int B, int N, int M, int D, int iterations,
                               float epsilon, float unused_score, int distance_l1, float* P, void* ws,
                               size_t ws_bytes, void* stream) {
    OM_ON_DEVICE_OF(desc1);
    return sinkhorn_launch(desc1, desc2, B, N, M, D, iterations, epsilon, unused_score, distance_l1, P, ws, ws_bytes,
                           (cudaStream_t)stream);
}
