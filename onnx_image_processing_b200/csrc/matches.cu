// Match extraction from the Sinkhorn probability matrix.
//
// Replaces (reference file:line):
//   matching/match_extraction.py:46-184   MutualNearestNeighborMatcher.forward
//
// P is (B, N+1, M+1); only the core block P[:, :N, :M] is looked at (:72).  Three small kernels on one stream:
//   row pass   one warp per row i: max_j P[i][j] and the first j that attains it (torch.argmax / max, :76-77)
//   col pass   one thread per column j: the first i that attains max_i P[i][j] (:80)
//   select     one CTA per pair: mutual check (:95-103), threshold (:106-109), score -1 for rejected rows
//              (:117-121), descending sort (torch.topk sorted, :124-130), zero padding when N < max_matches
//              (:133-142), keypoint gathers (:151-178), valid = score > 0 (:181)
// Equal scores are emitted in ascending row order (torch.topk leaves ties unspecified); the rows rejected by the
// mutual / threshold test all carry score -1 and therefore appear in ascending row order too.
#include <math_constants.h>

#include "common.cuh"

namespace om {

namespace {

constexpr int MT = 512;

__global__ void __launch_bounds__(MT) match_rows_kernel(const float* P, int N, int M, float* row_max, int* row_arg) {
    const int z = blockIdx.y;
    const int i = blockIdx.x * (MT / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= N) return;
    const float* row = P + ((size_t)z * (N + 1) + i) * (M + 1);
    float best = -CUDART_INF_F;
    int arg = 0x7fffffff;
    for (int j = lane; j < M; j += 32) {
        const float v = row[j];
        if (v > best || (v == best && j < arg) || arg == 0x7fffffff) { best = v; arg = j; }   // first maximal value
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (oa != 0x7fffffff && (arg == 0x7fffffff || ob > best || (ob == best && oa < arg))) { best = ob; arg = oa; }
    }
    if (lane == 0) {
        row_max[(size_t)z * N + i] = best;
        if (row_arg != nullptr) row_arg[(size_t)z * N + i] = arg;
    }
}

__global__ void __launch_bounds__(256) match_cols_kernel(const float* P, int N, int M, int* col_arg) {
    const int z = blockIdx.y;
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= M) return;
    const float* col = P + (size_t)z * (N + 1) * (M + 1) + j;
    float best = col[0];
    int arg = 0;
    for (int i = 1; i < N; ++i) {
        const float v = col[(size_t)i * (M + 1)];
        if (v > best) { best = v; arg = i; }                            // strict: the first maximal value wins
    }
    col_arg[(size_t)z * M + j] = arg;
}

// float -> unsigned with the same order (so that -1 sorts below every probability)
__device__ __forceinline__ unsigned int ordered_bits(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned int o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void __launch_bounds__(MT) match_select_kernel(const float* row_max, const int* row_arg, const int* col_arg,
                                                          const float* kpts1, const float* kpts2, int N, int M, int Np2,
                                                          int max_matches, float threshold, float* mk1, float* mk2,
                                                          float* scores, unsigned char* valid) {
    extern __shared__ unsigned long long keys[];   // Np2
    const int z = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < Np2; i += MT) {
        unsigned long long k = 0ull;               // padding sorts last
        if (i < N) {
            const int j = row_arg[(size_t)z * N + i];
            const float p = row_max[(size_t)z * N + i];
            const bool mutual = col_arg[(size_t)z * M + j] == i;                         // :95-103
            const float s = (mutual && p >= threshold) ? p : -1.0f;                      // :106-121
            k = ((unsigned long long)ordered_bits(s) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
        }
        keys[i] = k;
    }
    __syncthreads();
    for (int size = 2; size <= Np2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (Np2 >> 1); i += MT) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const unsigned long long x = keys[lo], y = keys[hi];
                if ((x < y) == desc) { keys[lo] = y; keys[hi] = x; }
            }
            __syncthreads();
        }
    }
    const int take = min(max_matches, N);
    for (int m = tid; m < max_matches; m += MT) {
        float s = 0.0f;                                                                  // :133-142 zero padding
        int i = 0;
        if (m < take) {
            const unsigned long long k = keys[m];
            s = from_ordered_bits((unsigned int)(k >> 32));
            i = (int)(0xFFFFFFFFu - (unsigned int)(k & 0xFFFFFFFFull));
        }
        const int ic = min(max(i, 0), N - 1);
        const int j = min(max(row_arg[(size_t)z * N + ic], 0), M - 1);                   // :164-172
        const size_t o = ((size_t)z * max_matches + m) * 2;
        mk1[o] = kpts1[((size_t)z * N + ic) * 2];
        mk1[o + 1] = kpts1[((size_t)z * N + ic) * 2 + 1];
        mk2[o] = kpts2[((size_t)z * M + j) * 2];
        mk2[o + 1] = kpts2[((size_t)z * M + j) * 2 + 1];
        scores[(size_t)z * max_matches + m] = s;
        valid[(size_t)z * max_matches + m] = s > 0.0f ? 1 : 0;                           // :181
    }
}

// SinkhornMatcherWithFilters epilogue, matching/sinkhorn.py:311-465, one warp per keypoint row i < N, in place on P:
//   ratio filter   best / (second + 1e-8) >= ratio_threshold over the core row (:332-346; topk(2): the two largest
//                  values, a duplicate maximum counts twice)
//   margin filter  best - P[i][M] >= dustbin_margin (:366-376)
//   rewrite        rejected rows: core -> 0, dustbin column -> 1 (:448-457); accepted rows unchanged
__global__ void __launch_bounds__(MT) filter_rows_kernel(float* P, int N, int M, float ratio_threshold, float dustbin_margin,
                                                         unsigned char* valid) {
    const int z = blockIdx.y;
    const int i = blockIdx.x * (MT / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= N) return;
    float* row = P + ((size_t)z * (N + 1) + i) * (M + 1);
    float b1 = -CUDART_INF_F, b2 = -CUDART_INF_F;                       // largest, second largest
    for (int j = lane; j < M; j += 32) {
        const float v = row[j];
        if (v > b1) { b2 = b1; b1 = v; } else if (v > b2) b2 = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float o1 = __shfl_xor_sync(0xffffffffu, b1, o), o2 = __shfl_xor_sync(0xffffffffu, b2, o);
        const float n1 = fmaxf(b1, o1);
        b2 = fmaxf(fminf(b1, o1), fmaxf(b2, o2));
        b1 = n1;
    }
    bool ok = true;
    if (ratio_threshold > 0.0f) {
        const float second = M >= 2 ? b2 : 0.0f;                        // :339-341
        ok = ok && (__fdiv_rn(b1, __fadd_rn(second, 1e-8f)) >= ratio_threshold);
    }
    if (dustbin_margin >= 0.0f) ok = ok && (__fsub_rn(b1, row[M]) >= dustbin_margin);
    if (!ok) {
        for (int j = lane; j < M; j += 32) row[j] = 0.0f;               // P * 0 (probabilities are finite and >= 0)
        if (lane == 0) row[M] = 1.0f;
    }
    if (lane == 0) valid[(size_t)z * N + i] = ok ? 1 : 0;
}

int next_pow2(int v) {
    int p = 2;
    while (p < v) p <<= 1;
    return p;
}

// SinkhornMatcherWithScores, matching/sinkhorn.py:251-257: column maxima of the core block (the row maxima come from
// match_rows_kernel).  One thread per column, consecutive threads read consecutive addresses of a row.
__global__ void __launch_bounds__(256) col_max_kernel(const float* P, int N, int M, float* col_max) {
    const int z = blockIdx.y;
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= M) return;
    const float* col = P + (size_t)z * (N + 1) * (M + 1) + j;
    float best = col[0];
    for (int i = 1; i < N; ++i) best = fmaxf(best, col[(size_t)i * (M + 1)]);
    col_max[(size_t)z * M + j] = best;
}

}  // namespace

size_t mutual_matches_workspace_bytes(int B, int N, int M) {
    if (B <= 0 || N <= 0 || M <= 0) return 0;
    return align_up((size_t)B * N * sizeof(float)) + align_up((size_t)B * N * sizeof(int)) + align_up((size_t)B * M * sizeof(int));
}

int filter_rows_launch(float* probs, int B, int N, int M, float ratio_threshold, float dustbin_margin, unsigned char* valid,
                       cudaStream_t st) {
    if (probs == nullptr || valid == nullptr) return OM_ERR_NULL;
    if (B <= 0 || N <= 0 || M <= 0) return OM_ERR_SHAPE;
    if (B > 65535) return OM_ERR_LIMIT;
    filter_rows_kernel<<<dim3((N + MT / 32 - 1) / (MT / 32), B), MT, 0, st>>>(probs, N, M, ratio_threshold, dustbin_margin, valid);
    OM_AFTER_LAUNCH();
    return OM_OK;
}

int sinkhorn_scores_launch(const float* probs, int B, int N, int M, float* scores0, float* scores1, cudaStream_t st) {
    if (probs == nullptr) return OM_ERR_NULL;
    if (B <= 0 || N <= 0 || M <= 0) return OM_ERR_SHAPE;
    if (B > 65535) return OM_ERR_LIMIT;
    if (scores0 != nullptr) {
        match_rows_kernel<<<dim3((N + MT / 32 - 1) / (MT / 32), B), MT, 0, st>>>(probs, N, M, scores0, nullptr);
        OM_AFTER_LAUNCH();
    }
    if (scores1 != nullptr) {
        col_max_kernel<<<dim3((M + 255) / 256, B), 256, 0, st>>>(probs, N, M, scores1);
        OM_AFTER_LAUNCH();
    }
    return OM_OK;
}

int mutual_matches_launch(const float* probs, const float* kpts1, const float* kpts2, int B, int N, int M, int max_matches,
                          float threshold, float* matched_kpts1, float* matched_kpts2, float* scores, unsigned char* valid,
                          void* ws, size_t ws_bytes, cudaStream_t st) {
    if (probs == nullptr || kpts1 == nullptr || kpts2 == nullptr || matched_kpts1 == nullptr || matched_kpts2 == nullptr ||
        scores == nullptr || valid == nullptr)
        return OM_ERR_NULL;
    if (B <= 0 || N <= 0 || M <= 0 || max_matches <= 0) return OM_ERR_SHAPE;
    if (B > 65535 || N > 16384) return OM_ERR_LIMIT;
    if (ws == nullptr || ws_bytes < mutual_matches_workspace_bytes(B, N, M)) return OM_ERR_WORKSPACE;
    char* p = (char*)ws;
    float* row_max = (float*)p;
    p += align_up((size_t)B * N * sizeof(float));
    int* row_arg = (int*)p;
    p += align_up((size_t)B * N * sizeof(int));
    int* col_arg = (int*)p;
    match_rows_kernel<<<dim3((N + MT / 32 - 1) / (MT / 32), B), MT, 0, st>>>(probs, N, M, row_max, row_arg);
    OM_AFTER_LAUNCH();
    match_cols_kernel<<<dim3((M + 255) / 256, B), 256, 0, st>>>(probs, N, M, col_arg);
    OM_AFTER_LAUNCH();
    const int Np2 = next_pow2(N);
    const size_t smem = (size_t)Np2 * sizeof(unsigned long long);
    OM_TRY(set_smem(match_select_kernel, smem));
    match_select_kernel<<<B, MT, smem, st>>>(row_max, row_arg, col_arg, kpts1, kpts2, N, M, Np2, max_matches, threshold,
                                             matched_kpts1, matched_kpts2, scores, valid);
    OM_AFTER_LAUNCH();
    return OM_OK;
}

}  // namespace om

using namespace om;

extern "C" int om_sinkhorn_filter_rows_f32(float* probs, int B, int N, int M, float ratio_threshold, float dustbin_margin,
                                           unsigned char* valid, void* stream) {
    OM_ON_DEVICE_OF(probs);
    return filter_rows_launch(probs, B, N, M, ratio_threshold, dustbin_margin, valid, (cudaStream_t)stream);
}

extern "C" int om_sinkhorn_scores_f32(const float* probs, int B, int N, int M, float* scores0, float* scores1, void* stream) {
    OM_ON_DEVICE_OF(probs);
    if (scores0 == nullptr || scores1 == nullptr) return OM_ERR_NULL;
    return sinkhorn_scores_launch(probs, B, N, M, scores0, scores1, (cudaStream_t)stream);
}

extern "C" size_t om_mutual_matches_workspace_bytes(int B, int N, int M) { return mutual_matches_workspace_bytes(B, N, M); }

extern "C" int om_mutual_matches_f32(const float* probs, const float* kpts1, const float* kpts2, int B, int N, int M,
                                     int max_matches, float threshold, float* matched_kpts1, float* matched_kpts2,
                                     float* scores, unsigned char* valid, void* ws, size_t ws_bytes, void* stream) {
    OM_ON_DEVICE_OF(probs);
    return mutual_matches_launch(probs, kpts1, kpts2, B, N, M, max_matches, threshold, matched_kpts1, matched_kpts2, scores, valid,
                                 ws, ws_bytes, (cudaStream_t)stream);
}
