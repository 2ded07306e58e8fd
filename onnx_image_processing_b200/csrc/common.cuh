// Shared host/device helpers for libom_b200.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/om_b200.h"

namespace om {

extern unsigned long long g_launches;  // defined in api.cu

inline void count_launch() { __atomic_add_fetch(&g_launches, 1ULL, __ATOMIC_RELAXED); }

#define OM_CUDA(expr)                                                      \
    do {                                                                   \
        cudaError_t e__ = (expr);                                          \
        if (e__ != cudaSuccess) return OM_ERR_CUDA_BASE + (int)e__;        \
    } while (0)

// after every <<<>>>: count the launch and surface configuration errors immediately
#define OM_AFTER_LAUNCH()                                                  \
    do {                                                                   \
        om::count_launch();                                                \
        cudaError_t e__ = cudaGetLastError();                              \
        if (e__ != cudaSuccess) return OM_ERR_CUDA_BASE + (int)e__;        \
    } while (0)

#define OM_TRY(expr)                                                       \
    do {                                                                   \
        int s__ = (expr);                                                  \
        if (s__ != OM_OK) return s__;                                      \
    } while (0)

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// implementation limits shared by the stage launchers and the fused matcher's up-front validation
constexpr int OM_MAX_BLOCK_HALF = 4;     // block_size <= 9
constexpr int OM_MAX_NMS_RADIUS = 8;
constexpr int OM_MAX_K = 16384;          // top-k kernel: keypoints per image

// Every compute entry point runs on the device that owns its first device pointer and restores the caller's current
// device before returning (the library links its own static CUDA runtime; a changed current device must not leak into
// the caller's).  One-GPU processes skip the pointer query.
int device_count_cached();
struct DeviceScope {
    int prev = -1, dev = 0, status = OM_OK;
    explicit DeviceScope(const void* device_ptr) {
        if (device_count_cached() <= 1 || device_ptr == nullptr) { cudaGetDevice(&dev); return; }
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, device_ptr) != cudaSuccess || at.type != cudaMemoryTypeDevice) {
            cudaGetLastError();
            cudaGetDevice(&dev);
            return;
        }
        dev = at.device;
        int cur = 0;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != dev) {
            if (cudaSetDevice(dev) == cudaSuccess) prev = cur;
            else status = OM_ERR_CUDA_BASE + (int)cudaGetLastError();
        }
    }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceScope(const DeviceScope&) = delete;
    DeviceScope& operator=(const DeviceScope&) = delete;
};
#define OM_ON_DEVICE_OF(ptr)                       \
    om::DeviceScope dev_scope__(ptr);              \
    if (dev_scope__.status != OM_OK) return dev_scope__.status

template <typename K>
inline int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return OM_ERR_CUDA_BASE + (int)e;
    }
    return OM_OK;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- mbarrier / TMA (cp.async.bulk.tensor) primitives ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA tile load global -> shared of one box of a rank-3 tensor map; completes `bytes of the box` on `bar`
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// Host: rank-3 tiled tensor map over a dense (d2, d1, d0) array of 4-byte (or 2-byte: elem_bytes) elements whose rows are
// `pitch_elems` apart (whole 16-byte pieces, base 16-byte aligned); box = (1, box1, box0), no swizzle,
// out-of-bounds elements read as zero.  Defined in api.cu (driver entry point resolved at run time).
int make_tmap_3d(CUtensorMap* map, bool is_float, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t pitch_elems, uint32_t box0, uint32_t box1, int elem_bytes = 4);

// ---- internal stage launchers shared between the per-stage C ABI and the fused matcher -------
struct DetectCfg {
    int B, H, W;
    int block_size, nms_radius, border_margin;
    float score_threshold;
    int K;
    int image_u8;          // 1: the image pointer is uint8, 0: float32
};

size_t topk_workspace_bytes(int B, int H, int W, int K);
int detect_launch(const void* image, const DetectCfg& c, float* score_map, float* kpts, float* kpt_scores,
                  void* ws, size_t ws_bytes, cudaStream_t st);

// NMS + top-k on a caller's score map (cfg.block_size / image_u8 unused)
int detect_scores_launch(const float* scores, const DetectCfg& c, float* kpts, float* kpt_scores, void* ws, size_t ws_bytes,
                         cudaStream_t st);

size_t sparse_bad_workspace_bytes(int B, int H, int W, int theta_mode);
int sparse_bad_launch(const void* image, int image_u8, int B, int H, int W, const float* kpts, int K, const float* pair_table,
                      int P, int desc_mode, float temperature, int normalize, int sampling_mode, int theta_mode,
                      const float* orientation, const float* moment_kernels, int patch_size, float* desc,
                      void* ws, size_t ws_bytes, cudaStream_t st, int phase = 0);   // phase 1: integral only, 2: descriptors only

size_t dense_bad_workspace_bytes(int B, int H, int W);
int dense_bad_at_kpts_launch(const void* image, int image_u8, int B, int H, int W, const float* kpts, int K,
                             const float* pair_table, int P, int desc_mode, float temperature, int normalize,
                             float* desc, void* ws, size_t ws_bytes, cudaStream_t st, int phase = 0);

// What the Sinkhorn stage can produce besides (or instead of) the probability matrix: fused into the tcgen05 cluster kernel's
// epilogue when that kernel runs in its scaling form, separate kernels on a stored P otherwise.
struct SinkhornEpilogue {
    float* scores0 = nullptr;             // (B,N) best core probability per row        SinkhornMatcherWithScores, sinkhorn.py:251-257
    float* scores1 = nullptr;             // (B,M) best core probability per column
    int filters = 0;                      // SinkhornMatcherWithFilters, sinkhorn.py:311-465
    float ratio_threshold = -1.0f, dustbin_margin = -1.0f;
    unsigned char* filter_valid = nullptr;   // (B,N)
    int matches = 0;                      // MutualNearestNeighborMatcher, match_extraction.py:46-184
    const float* kpts1 = nullptr;
    const float* kpts2 = nullptr;
    int max_matches = 0;
    float match_threshold = 0.0f;
    float* mk1 = nullptr;                 // (B,max_matches,2)
    float* mk2 = nullptr;
    float* mscores = nullptr;             // (B,max_matches)
    unsigned char* mvalid = nullptr;      // (B,max_matches)
    bool any() const { return scores0 != nullptr || scores1 != nullptr || filters != 0 || matches != 0; }
};

extern int g_tc_allow_scaling, g_tc_allow_f16;   // sinkhorn_tc.cu; test hooks
// tcgen05 / TMEM cluster kernel (sinkhorn_tc.cu); limits: L2 cost, N <= 512, M <= 512, D % 16 == 0
int sinkhorn_cluster_tc(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float eps,
                        float unused, float* P, cudaStream_t st);
// the same with the fused epilogue; P may be null.  Returns OM_ERR_PARAM when this configuration cannot take the fused form
// (log-domain loop needed, descriptor length not a multiple of 32): the caller then stores P and runs the separate kernels.
int sinkhorn_cluster_tc_epi(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float eps,
                            float unused, float* P, const SinkhornEpilogue& e, cudaStream_t st);
bool sinkhorn_epilogue_can_fuse(int N, int M, int D, float eps, float unused, int distance_l1);
// hybrid-resident cluster kernel (sinkhorn_hy.cu): K half in registers, half in shared memory; 4-CTA clusters up to
// 512 x 512, 16-CTA clusters up to 1024 x 1024; scaling form, squared-L2 cost, D % 32 == 0.  `e` may be null (P only).
bool sinkhorn_hy_eligible(int N, int M, int D, float eps, float unused, int distance_l1);
size_t sinkhorn_hy_workspace_bytes(int B, int N, int M, int D);
int sinkhorn_hy_launch(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float eps, float unused,
                       float* P, const SinkhornEpilogue* e, void* ws, size_t ws_bytes, cudaStream_t st);
// the three steps of sinkhorn_hy_launch, and the router's answer to "will sinkhorn_launch / sinkhorn_ex_launch take the hybrid
// kernel for this problem?" (the fused matcher then packs each image's descriptors on that image's stream)
int sinkhorn_hy_prepare(int B, int N, int M, int D, void* ws, size_t ws_bytes, cudaStream_t st);
// binary = 1: the rows are known to hold only 0 and one common value (hard-binarised BAD descriptors): one 8-bit operand
// term, the similarity GEMM as popcounts on tcgen05 kind::f8f6f4 (pack and run must agree; sinkhorn_hy_binary_ok says whether
// the descriptor length allows it)
bool sinkhorn_hy_binary_ok(int N, int M, int D);
int sinkhorn_hy_pack(int which, const float* d, int B, int N, int M, int D, void* ws, size_t ws_bytes, cudaStream_t st, int binary = 0);
int sinkhorn_hy_run(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float eps, float unused,
                    float* P, const SinkhornEpilogue* e, void* ws, size_t ws_bytes, cudaStream_t st, int binary = 0);
bool sinkhorn_routes_to_hy(int B, int N, int M, int D, float epsilon, float unused_score, int distance_l1, const void* ws,
                           size_t ws_bytes);
// streaming kernels beyond the hybrid kernel's sizes (sinkhorn_xl.cu): the same three steps
bool sinkhorn_routes_to_xl(int B, int N, int M, int D, float epsilon, float unused_score, int distance_l1, const void* ws,
                           size_t ws_bytes);
int sinkhorn_xl_prepare(int B, int N, int M, int D, void* ws, size_t ws_bytes, cudaStream_t st);
int sinkhorn_xl_pack(int which, const float* d, int B, int N, int M, int D, void* ws, size_t ws_bytes, cudaStream_t st);
int sinkhorn_xl_run(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float eps, float unused,
                    float* P, void* ws, size_t ws_bytes, cudaStream_t st);
size_t sinkhorn_ex_workspace_bytes(int B, int N, int M, int D);
// Sinkhorn + optional outputs; P may be null when the epilogue is fused (otherwise it is kept in the workspace)
int sinkhorn_ex_launch(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float epsilon,
                       float unused_score, int distance_l1, float* P, const SinkhornEpilogue& e, void* ws, size_t ws_bytes,
                       cudaStream_t st);
SinkhornEpilogue epilogue_from_outputs(const om_sinkhorn_outputs* o);
// separate-kernel forms of the epilogue pieces (matches.cu)
int mutual_matches_launch(const float* probs, const float* kpts1, const float* kpts2, int B, int N, int M, int max_matches,
                          float threshold, float* mk1, float* mk2, float* scores, unsigned char* valid, void* ws, size_t ws_bytes,
                          cudaStream_t st);
size_t mutual_matches_workspace_bytes(int B, int N, int M);
int filter_rows_launch(float* probs, int B, int N, int M, float ratio_threshold, float dustbin_margin, unsigned char* valid,
                       cudaStream_t st);
int sinkhorn_scores_launch(const float* probs, int B, int N, int M, float* scores0, float* scores1, cudaStream_t st);

size_t sinkhorn_workspace_bytes(int B, int N, int M, int D);
// cost matrix of the generic Sinkhorn path on tcgen05 (sinkhorn_tc.cu); D % 32 == 0; *ovf != 0: does nothing
int cost_tc_launch(const float* d1, const float* d2, const float* n1, const float* n2, int B, int N, int M, int D, float eps,
                   float dustbin, int as_exp, float* S, const unsigned int* ovf, cudaStream_t st);
int sinkhorn_launch(const float* d1, const float* d2, int B, int N, int M, int D, int iterations, float epsilon,
                    float unused_score, int distance_l1, float* P, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace om
