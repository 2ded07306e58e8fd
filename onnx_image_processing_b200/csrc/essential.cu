// Essential-matrix head on the Sinkhorn probability matrix, batched over pairs.
//
// Replaces (reference file:line):
//   geometry/essential_matrix_estimator.py:292-392                       EssentialMatrixEstimator.forward (grid points)
//   feature_detection/shi_tomasi_angle_sparse_bad_sinkhorn_essential_matrix.py:184-271   _estimate_essential_matrix
//   helpers: _hartley_normalization :242-290, _min_eigvec9 :150-172, _project_onto_E_manifold :174-240, _det3 :126-130
//
// One CTA per pair; the reference handles one pair per call.  P is (B, N+1, M+1), only the core block is used.
//   A  thresholds   k-th largest masked probability of every row (one warp per row) and column (one thread per
//                   column), torch.topk semantics: duplicates count (:221-233)
//   B  weights      w = P if it passes both thresholds and P > 0.01, else 0 (:235-237); marginals w1, w2 (:240-241)
//   C  Hartley      weighted centroid and RMS distance of both point sets (:243-247)
//   D  normal eq.   M_flat[pr][qs] = sum_ij w_ij f1_i[p] f1_i[r] f2_j[q] f2_j[s]  (:252-262), index permutation :263
//   E  algebra      30 shifted power iterations on the 9x9, denormalisation, projection onto singular values (s, s, 0)
// The long sums (marginals, centroids, the 81 entries of M_flat) are accumulated in double in a fixed order and rounded
// to float once; the reference accumulates them in float inside ATen's sum / matmul (order unspecified), so the two
// agree to float rounding of those sums.  Phase E repeats the reference's float operations in index order.
#include <cooperative_groups.h>
#include <math_constants.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace om {

namespace {

constexpr int ET = 512;              // threads per CTA
constexpr int EW = ET / 32;          // warps
constexpr int E_MAXK = 8;            // largest supported top_k

struct EssArgs {
    const float* P;                  // (B, N+1, M+1)
    const float* pts1;               // (B or 1, N, 2) normalised (x, y)
    const float* pts2;               // (B or 1, M, 2)
    const unsigned char* valid1;     // (B, N) or null
    const unsigned char* valid2;     // (B, M) or null
    int N, M, top_k, n_iter, n_iter_manifold;
    int pts_batched;                 // 0: one point set shared by all pairs (grid form)
    float* E;                        // (B, 3, 3)
};

// descending insertion into a register list of the k largest values seen so far
template <int MAXK>
__device__ __forceinline__ void topk_insert(float (&t)[MAXK], int k, float v) {
    if (!(v > t[k - 1])) return;
#pragma unroll
    for (int a = MAXK - 1; a >= 0; --a) {
        if (a < k) {
            const float up = a > 0 ? t[a - 1] : CUDART_INF_F;
            if (v > up) t[a] = up;                   // shift down
            else if (v > t[a]) t[a] = v;             // lands here (every slot below was shifted already)
        }
    }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum of one double per thread over the CTA, in a fixed order; result in every thread
__device__ double block_sum_d(double v, double* scratch) {
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) scratch[wrp] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < EW; ++w) s += scratch[w];
    return s;
}

struct Hartley {
    float s, cx, cy;
};

// essential_matrix_estimator.py:264-274 from the three weighted sums (float operations in the reference's order)
__device__ Hartley hartley_from(const float* pts, const float* w, int n, double* scratch) {
    double sw = 0.0, sx = 0.0, sy = 0.0;
    for (int i = threadIdx.x; i < n; i += ET) {
        const float wi = w[i];
        sw += (double)wi;
        sx += (double)__fmul_rn(wi, pts[2 * i]);
        sy += (double)__fmul_rn(wi, pts[2 * i + 1]);
    }
    const float w_sum = __fadd_rn((float)block_sum_d(sw, scratch), 1e-8f);
    Hartley h;
    h.cx = __fdiv_rn((float)block_sum_d(sx, scratch), w_sum);
    h.cy = __fdiv_rn((float)block_sum_d(sy, scratch), w_sum);
    double sd = 0.0;
    for (int i = threadIdx.x; i < n; i += ET) {
        const float dx = __fsub_rn(pts[2 * i], h.cx), dy = __fsub_rn(pts[2 * i + 1], h.cy);
        const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        sd += (double)__fmul_rn(w[i], d2);
    }
    const float mean_dist = sqrtf(__fadd_rn(__fdiv_rn((float)block_sum_d(sd, scratch), w_sum), 1e-8f));
    h.s = __fdiv_rn(sqrtf(2.0f), __fadd_rn(mean_dist, 1e-8f));
    return h;
}

// ---- 3x3 helpers for phase E (one thread) ----------------------------------------------------
struct M3 {
    float a[3][3];
};
__device__ M3 mul3(const M3& x, const M3& y) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float s = 0.0f;
            for (int k = 0; k < 3; ++k) s = __fmaf_rn(x.a[i][k], y.a[k][j], s);
            r.a[i][j] = s;
        }
    return r;
}
__device__ M3 transpose3(const M3& x) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.a[i][j] = x.a[j][i];
    return r;
}
__device__ void matvec3(const M3& m, const float (&v)[3], float (&o)[3]) {
    for (int i = 0; i < 3; ++i) {
        float s = 0.0f;
        for (int k = 0; k < 3; ++k) s = __fmaf_rn(m.a[i][k], v[k], s);
        o[i] = s;
    }
}
__device__ float norm3(const float (&v)[3]) { return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
__device__ void cross3(const float (&a)[3], const float (&b)[3], float (&o)[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ float det3(const M3& m) {                 // essential_matrix_estimator.py:126-130
    return m.a[0][0] * (m.a[1][1] * m.a[2][2] - m.a[1][2] * m.a[2][1]) - m.a[0][1] * (m.a[1][0] * m.a[2][2] - m.a[1][2] * m.a[2][0]) +
           m.a[0][2] * (m.a[1][0] * m.a[2][1] - m.a[1][1] * m.a[2][0]);
}
__device__ float sign1(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
__device__ void power3(const M3& b, int iters, float (&v)[3]) {
    v[0] = v[1] = v[2] = __fdiv_rn(1.0f, sqrtf(3.0f));
    for (int it = 0; it < iters; ++it) {
        float t[3];
        matvec3(b, v, t);
        const float d = __fadd_rn(norm3(t), 1e-8f);
        for (int k = 0; k < 3; ++k) v[k] = __fdiv_rn(t[k], d);
    }
}

// essential_matrix_estimator.py:174-240
__device__ M3 project_onto_manifold(const M3& E, int iters) {
    const M3 B = mul3(transpose3(E), E);
    const float lam = (B.a[0][0] + B.a[1][1]) + B.a[2][2];
    float v1[3], v3[3], v2[3];
    power3(B, iters, v1);
    M3 Bs;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Bs.a[i][j] = (i == j ? lam : 0.0f) - B.a[i][j];
    power3(Bs, iters, v3);
    cross3(v3, v1, v2);
    {
        const float d = __fadd_rn(norm3(v2), 1e-8f);
        for (int k = 0; k < 3; ++k) v2[k] = __fdiv_rn(v2[k], d);
    }
    M3 V;
    for (int i = 0; i < 3; ++i) { V.a[i][0] = v1[i]; V.a[i][1] = v2[i]; V.a[i][2] = v3[i]; }
    const float sv = sign1(det3(V));
    for (int i = 0; i < 3; ++i) V.a[i][2] *= sv;
    float c0[3] = {V.a[0][0], V.a[1][0], V.a[2][0]}, c1[3] = {V.a[0][1], V.a[1][1], V.a[2][1]}, e0[3], e1[3];
    matvec3(E, c0, e0);
    matvec3(E, c1, e1);
    const float sigma1 = norm3(e0), sigma2 = norm3(e1);
    const float s_avg = __fdiv_rn(__fadd_rn(sigma1, sigma2), 2.0f);
    float u1[3], u2[3], u3[3];
    for (int k = 0; k < 3; ++k) {
        u1[k] = __fdiv_rn(e0[k], __fadd_rn(sigma1, 1e-8f));
        u2[k] = __fdiv_rn(e1[k], __fadd_rn(sigma2, 1e-8f));
    }
    cross3(u1, u2, u3);
    M3 U;
    for (int i = 0; i < 3; ++i) { U.a[i][0] = u1[i]; U.a[i][1] = u2[i]; U.a[i][2] = u3[i]; }
    const float su = sign1(det3(U));
    for (int i = 0; i < 3; ++i) U.a[i][2] *= su;
    M3 S;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) S.a[i][j] = (i == j && i < 2) ? s_avg : 0.0f;
    return mul3(mul3(U, S), transpose3(V));
}

// CLS = CTAs per pair.  CLS = 8: a thread-block cluster shares one pair; CTA r owns the rows [r * RS, (r + 1) * RS) of P for
// every pass (row statistics are complete per CTA, column statistics are partial per CTA and combined through
// distributed shared memory after a cluster barrier), rank 0 finishes the algebra.  CLS = 1: one CTA does everything
// (used when the per-column scratch does not fit shared memory).  One pair on one SM took 0.65 ms (2.7 M warp
// instructions on 16 warps); eight SMs per pair bring a single pair to ~0.1 ms.
template <int CLS>
__global__ void __launch_bounds__(ET) essential_kernel(EssArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double scratch[EW];
    __shared__ double mpart[EW][81];
    __shared__ double mcta[81];
    __shared__ float wf2[EW][6];
    __shared__ float m9[81];
    const int N = a.N, M = a.M, k = a.top_k;
    float* thr_row = reinterpret_cast<float*>(smem_raw);
    float* thr_col = thr_row + N;
    float* w1 = thr_col + M;
    float* w2 = w1 + N;
    float* mul1 = w2 + M;                            // validity factors (1 or 0)
    float* mul2 = mul1 + N;
    double* w2part = reinterpret_cast<double*>(mul2 + M + ((3 * N + 3 * M) & 1));   // CLS > 1: column sums over the own rows
    float* colk = reinterpret_cast<float*>(w2part + M);                             // CLS > 1: k largest of each column over the own rows
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    int rank = 0;
    if constexpr (CLS > 1) rank = (int)cg::this_cluster().block_rank();
    const int z = blockIdx.x / CLS;
    const int RS = (N + CLS - 1) / CLS;              // rows per CTA
    const int i_lo = min(rank * RS, N), i_hi = min(i_lo + RS, N);
    const float* P = a.P + (size_t)z * (N + 1) * (M + 1);
    const float* pts1 = a.pts1 + (a.pts_batched ? (size_t)z * N * 2 : 0);
    const float* pts2 = a.pts2 + (a.pts_batched ? (size_t)z * M * 2 : 0);
    const unsigned full = 0xffffffffu;
    auto csync = [&]() {
        if constexpr (CLS > 1) cg::this_cluster().sync();
        else __syncthreads();
    };

    for (int i = tid; i < N; i += ET) mul1[i] = (a.valid1 == nullptr || a.valid1[(size_t)z * N + i]) ? 1.0f : 0.0f;
    for (int j = tid; j < M; j += ET) mul2[j] = (a.valid2 == nullptr || a.valid2[(size_t)z * M + j]) ? 1.0f : 0.0f;
    __syncthreads();
    auto masked = [&](int i, int j) -> float {       // ..._essential_matrix.py:216-218
        return __fmul_rn(__fmul_rn(P[(size_t)i * (M + 1) + j], mul1[i]), mul2[j]);
    };

    // ---- A: k-th largest of every row and column ------------------------------------------------------------------
    for (int i = i_lo + wrp; i < i_hi; i += EW) {
        float t[E_MAXK];
#pragma unroll
        for (int q = 0; q < E_MAXK; ++q) t[q] = -CUDART_INF_F;
        for (int j = lane; j < M; j += 32) topk_insert(t, k, masked(i, j));
        float kth = -CUDART_INF_F;
        for (int r = 0; r < k; ++r) {                // pop the warp-wide maximum k times
            float m = t[0];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(full, m, o));
            kth = m;
            const unsigned owners = __ballot_sync(full, t[0] == m);
            if (lane == __ffs(owners) - 1) {
#pragma unroll
                for (int q = 0; q < E_MAXK - 1; ++q) t[q] = t[q + 1];
                t[E_MAXK - 1] = -CUDART_INF_F;
            }
        }
        if (lane == 0) thr_row[i] = kth;
    }
    for (int j = tid; j < M; j += ET) {
        float t[E_MAXK];
#pragma unroll
        for (int q = 0; q < E_MAXK; ++q) t[q] = -CUDART_INF_F;
        for (int i = i_lo; i < i_hi; ++i) topk_insert(t, k, masked(i, j));
        if constexpr (CLS > 1) {
#pragma unroll
            for (int q = 0; q < E_MAXK; ++q)
                if (q < k) colk[(size_t)j * k + q] = t[q];
        } else {
            float kth = t[0];
#pragma unroll
            for (int q = 1; q < E_MAXK; ++q)
                if (q < k) kth = t[q];
            thr_col[j] = kth;
        }
    }
    csync();
    if constexpr (CLS > 1) {                         // merge the CLS partial lists of every column (each CTA, all columns)
        for (int j = tid; j < M; j += ET) {
            float t[E_MAXK];
#pragma unroll
            for (int q = 0; q < E_MAXK; ++q) t[q] = -CUDART_INF_F;
            for (int rr = 0; rr < CLS; ++rr) {
                const float* rk = cg::this_cluster().map_shared_rank(colk, rr);
                for (int q = 0; q < k; ++q) topk_insert(t, k, rk[(size_t)j * k + q]);
            }
            float kth = t[0];
#pragma unroll
            for (int q = 1; q < E_MAXK; ++q)
                if (q < k) kth = t[q];
            thr_col[j] = kth;
        }
        __syncthreads();
    }
    auto weight = [&](int i, int j) -> float {       // :221-237
        const float v = masked(i, j);
        return (v >= thr_row[i] && v >= thr_col[j] && v > 0.01f) ? v : 0.0f;
    };

    // ---- B: marginals ----------------------------------------------------------------------------------------------
    for (int i = i_lo + wrp; i < i_hi; i += EW) {
        double s = 0.0;
        for (int j = lane; j < M; j += 32) s += (double)weight(i, j);
        s = warp_sum_d(s);
        if (lane == 0) w1[i] = (float)s;
    }
    for (int j = tid; j < M; j += ET) {
        double s = 0.0;
        for (int i = i_lo; i < i_hi; ++i) s += (double)weight(i, j);
        if constexpr (CLS > 1) w2part[j] = s;
        else w2[j] = (float)s;
    }
    csync();
    if constexpr (CLS > 1) {                         // column sums over all rows; the other CTAs' slices of the row sums
        for (int j = tid; j < M; j += ET) {
            double s = 0.0;
            for (int rr = 0; rr < CLS; ++rr) s += cg::this_cluster().map_shared_rank(w2part, rr)[j];
            w2[j] = (float)s;
        }
        for (int i = tid; i < N; i += ET)
            if (i < i_lo || i >= i_hi) w1[i] = cg::this_cluster().map_shared_rank(w1, i / RS)[i];
        __syncthreads();
    }

    // ---- C: Hartley normalisation of both point sets (every CTA, from the complete marginals) --------------------
    const Hartley h1 = hartley_from(pts1, w1, N, scratch);
    const Hartley h2 = hartley_from(pts2, w2, M, scratch);

    // ---- D: M_flat = F1^T (W F2), one warp per row of W ------------------------------------------------------------
    double acc[3] = {0.0, 0.0, 0.0};                 // entries lane, lane + 32, lane + 64 of the 9x9 (pr, qs)
    for (int i = i_lo + wrp; i < i_hi; i += EW) {
        if (w1[i] == 0.0f) continue;                 // every weight of the row is zero (weights are >= 0)
        double s[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // sum_j w f2 f2^T: xx, xy, x, yy, y, 1
        for (int j = lane; j < M; j += 32) {
            const float w = weight(i, j);
            if (w != 0.0f) {
                const float x = __fmul_rn(__fsub_rn(pts2[2 * j], h2.cx), h2.s), y = __fmul_rn(__fsub_rn(pts2[2 * j + 1], h2.cy), h2.s);
                const double wd = (double)w;
                s[0] += wd * (double)__fmul_rn(x, x);
                s[1] += wd * (double)__fmul_rn(x, y);
                s[2] += wd * (double)x;
                s[3] += wd * (double)__fmul_rn(y, y);
                s[4] += wd * (double)y;
                s[5] += wd;
            }
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) s[q] = warp_sum_d(s[q]);
        __syncwarp();
        if (lane < 6) wf2[wrp][lane] = (float)(lane == 0 ? s[0] : lane == 1 ? s[1] : lane == 2 ? s[2] : lane == 3 ? s[3] : lane == 4 ? s[4] : s[5]);
        __syncwarp();
        const float x = __fmul_rn(__fsub_rn(pts1[2 * i], h1.cx), h1.s), y = __fmul_rn(__fsub_rn(pts1[2 * i + 1], h1.cy), h1.s);
        const float f1[3] = {x, y, 1.0f};
#pragma unroll
        for (int e3 = 0; e3 < 3; ++e3) {
            const int e = lane + 32 * e3;
            if (e < 81) {
                const int pr = e / 9, qs = e - 9 * pr;
                const int p = pr / 3, r = pr - 3 * p, q = qs / 3, sidx = qs - 3 * q;
                const int sym = q <= sidx ? (q == 0 ? sidx : q == 1 ? 2 + sidx : 5) : (sidx == 0 ? q : sidx == 1 ? 2 + q : 5);
                acc[e3] += (double)__fmul_rn(f1[p], f1[r]) * (double)wf2[wrp][sym];
            }
        }
    }
#pragma unroll
    for (int e3 = 0; e3 < 3; ++e3) {
        const int e = lane + 32 * e3;
        if (e < 81) mpart[wrp][e] = acc[e3];
    }
    __syncthreads();
    if (tid < 81) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < EW; ++w) s += mpart[w][tid];
        mcta[tid] = s;
    }
    csync();
    if (rank == 0 && tid < 81) {
        double s = 0.0;
        if constexpr (CLS > 1) {
            for (int rr = 0; rr < CLS; ++rr) s += cg::this_cluster().map_shared_rank(mcta, rr)[tid];
        } else {
            s = mcta[tid];
        }
        // (pr, qs) -> (pq, rs): M_mat[3p+q][3r+s] = M_flat[3p+r][3q+s]   (:263)
        const int pr = tid / 9, qs = tid - 9 * pr;
        const int p = pr / 3, r = pr - 3 * p, q = qs / 3, sidx = qs - 3 * q;
        m9[(3 * p + q) * 9 + (3 * r + sidx)] = (float)s;
    }
    csync();                                         // rank 0 has read every CTA's partial sums: the others may leave
    if (rank != 0) return;

    // ---- E: eigenvector, denormalisation, manifold projection (warp 0) --------------------------------------------
    if (wrp != 0) return;
    float lam = 0.0f;
    for (int d = 0; d < 9; ++d) lam = __fadd_rn(lam, m9[d * 9 + d]);
    float row[9];
    const int rr = lane < 9 ? lane : 0;
#pragma unroll
    for (int c = 0; c < 9; ++c) row[c] = __fsub_rn(c == rr ? lam : 0.0f, m9[rr * 9 + c]);   // trace(M) I - M  (:163-164)
    float v = __fdiv_rn(1.0f, 3.0f);
    for (int it = 0; it < a.n_iter; ++it) {          // :167-171
        float dot = 0.0f;
#pragma unroll
        for (int c = 0; c < 9; ++c) dot = __fmaf_rn(row[c], __shfl_sync(full, v, c), dot);
        float n2 = 0.0f;
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            const float t = __shfl_sync(full, dot, c);
            n2 = __fmaf_rn(t, t, n2);
        }
        v = __fdiv_rn(dot, __fadd_rn(sqrtf(n2), 1e-8f));
    }
    M3 Eraw;
#pragma unroll
    for (int c = 0; c < 9; ++c) Eraw.a[c / 3][c % 3] = __shfl_sync(full, v, c);
    if (lane != 0) return;
    M3 T1, T2;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T1.a[i][j] = T2.a[i][j] = 0.0f;
    T1.a[0][0] = T1.a[1][1] = h1.s; T1.a[0][2] = __fmul_rn(-h1.s, h1.cx); T1.a[1][2] = __fmul_rn(-h1.s, h1.cy); T1.a[2][2] = 1.0f;
    T2.a[0][0] = T2.a[1][1] = h2.s; T2.a[0][2] = __fmul_rn(-h2.s, h2.cx); T2.a[1][2] = __fmul_rn(-h2.s, h2.cy); T2.a[2][2] = 1.0f;
    const M3 Ed = mul3(mul3(transpose3(T2), Eraw), T1);                                     // :259
    const M3 Eo = project_onto_manifold(Ed, a.n_iter_manifold);
    float* out = a.E + (size_t)z * 9;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) out[i * 3 + j] = Eo.a[i][j];
}

int g_essential_cluster = 1;   // om_debug_essential_variant: 0 = one CTA per pair, 1 = an 8-CTA cluster per pair (default)

}  // namespace

}  // namespace om

using namespace om;

extern "C" void om_debug_essential_variant(int clustered) { g_essential_cluster = clustered ? 1 : 0; }

extern "C" int om_essential_matrix_f32(const float* probs, const float* pts1, const float* pts2, const unsigned char* valid1,
                                       const unsigned char* valid2, int B, int N, int M, int pts_batched, int top_k,
                                       int n_iter, int n_iter_manifold, float* E, void* stream) {
    OM_ON_DEVICE_OF(probs);
    if (probs == nullptr || pts1 == nullptr || pts2 == nullptr || E == nullptr) return OM_ERR_NULL;
    if ((valid1 == nullptr) != (valid2 == nullptr)) return OM_ERR_NULL;
    if (B <= 0 || N <= 0 || M <= 0) return OM_ERR_SHAPE;
    if (top_k < 1 || top_k > N || top_k > M) return OM_ERR_SHAPE;         // torch.topk raises for k > dim
    if (n_iter < 0 || n_iter_manifold < 0) return OM_ERR_PARAM;
    if (top_k > E_MAXK || N > 8192 || M > 8192) return OM_ERR_LIMIT;
    EssArgs a{probs, pts1, pts2, valid1, valid2, N, M, top_k, n_iter, n_iter_manifold, pts_batched, E};
    // floats: thresholds, marginals, validity factors (3N + 3M); cluster form: + one double and top_k floats per column
    const size_t base = (size_t)(3 * N + 3 * M + ((3 * N + 3 * M) & 1)) * sizeof(float);
    const size_t clustered = base + (size_t)M * sizeof(double) + (size_t)M * top_k * sizeof(float);
    constexpr int CLS = 8;
    // many small pairs already fill the GPU with one CTA each (measured: 64 pairs of 128 points 98 us vs 154 us clustered)
    const bool small_and_many = B >= 48 && (long long)N * M <= 256ll * 256ll;
    if (g_essential_cluster && !small_and_many && clustered + 16 * 1024 <= 200 * 1024 && (long long)B * CLS < (1ll << 31)) {
        if (clustered > 32 * 1024)
            OM_CUDA(cudaFuncSetAttribute(essential_kernel<CLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)clustered));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(B * CLS), 1, 1);
        cfg.blockDim = dim3(ET, 1, 1);
        cfg.dynamicSmemBytes = clustered;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CLS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        OM_CUDA(cudaLaunchKernelEx(&cfg, essential_kernel<CLS>, a));
    } else {
        // the kernel also has ~12 KB of static shared memory: opt in whenever the sum may pass the 48 KB default
        if (base > 32 * 1024) OM_CUDA(cudaFuncSetAttribute(essential_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)base));
        essential_kernel<1><<<B, ET, base, (cudaStream_t)stream>>>(a);
    }
    OM_AFTER_LAUNCH();
    return OM_OK;
}
