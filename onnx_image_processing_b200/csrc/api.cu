// Library-level entry points: version, error strings, launch counter, BAD tables, and the fused
// matcher that chains the stage launchers on one stream (no host synchronisation in between).
#include <string.h>

#include <mutex>

#include "bad_tables.inc"
#include "common.cuh"

namespace om {
unsigned long long g_launches = 0;

// cuTensorMapEncodeTiled lives in libcuda; the library links the runtime statically and does not link the
// driver, so the entry point is resolved through the runtime the first time a tensor map is needed.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_3d(CUtensorMap* map, bool is_float, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t pitch_elems, uint32_t box0, uint32_t box1, int elem_bytes) {
    static EncodeTiledFn encode = nullptr;
    if (encode == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        OM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (fn == nullptr || qres != cudaDriverEntryPointSuccess) return OM_ERR_CUDA_BASE + (int)cudaErrorNotSupported;
        encode = (EncodeTiledFn)fn;
    }
    if (elem_bytes != 4 && elem_bytes != 2) return OM_ERR_PARAM;
    const uint64_t per16 = 16 / (uint64_t)elem_bytes;                          // rows and boxes are whole 16-byte pieces
    if (pitch_elems % per16 != 0 || ((uintptr_t)base & 15) != 0 || box0 % per16 != 0 || box0 > 256 || box1 > 256) return OM_ERR_PARAM;
    const cuuint64_t dims[3] = {d0, d1, d2};
    const cuuint64_t strides[2] = {pitch_elems * elem_bytes, pitch_elems * elem_bytes * d1};     // bytes, dims 1 and 2
    const cuuint32_t box[3] = {box0, box1, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : (is_float ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT32), 3,
                              const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? OM_OK : OM_ERR_CUDA_BASE + (int)cudaErrorInvalidValue;
}
}  // namespace om

using namespace om;

extern "C" int om_version(void) { return 100; }

extern "C" int om_set_device(int device) {
    OM_CUDA(cudaSetDevice(device));
    return OM_OK;
}

extern "C" unsigned long long om_launch_count(void) { return __atomic_load_n(&om::g_launches, __ATOMIC_RELAXED); }

extern "C" const char* om_error_string(int status) {
    switch (status) {
        case OM_OK: return "ok";
        case OM_ERR_NULL: return "required pointer is NULL";
        case OM_ERR_SHAPE: return "invalid or inconsistent sizes";
        case OM_ERR_PARAM: return "unsupported parameter value";
        case OM_ERR_WORKSPACE: return "workspace missing or too small";
        case OM_ERR_LIMIT: return "size beyond an implementation limit";
        default: break;
    }
    if (status >= OM_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(status - OM_ERR_CUDA_BASE));
    return "unknown status";
}

extern "C" int om_bad_table(int num_pairs, signed char* h_boxes, float* h_thresholds) {
    if (h_boxes == nullptr || h_thresholds == nullptr) return OM_ERR_NULL;
    if (num_pairs == 256) {
        memcpy(h_boxes, OM_BAD_BOX_256, sizeof(OM_BAD_BOX_256));
        memcpy(h_thresholds, OM_BAD_THR_256, sizeof(OM_BAD_THR_256));
        return OM_OK;
    }
    if (num_pairs == 512) {
        memcpy(h_boxes, OM_BAD_BOX_512, sizeof(OM_BAD_BOX_512));
        memcpy(h_thresholds, OM_BAD_THR_512, sizeof(OM_BAD_THR_512));
        return OM_OK;
    }
    return OM_ERR_PARAM;   // descriptor/bad_params.py:1565 raises ValueError
}

namespace {

struct MatchWs {
    void* detect[2];  size_t detect_bytes;     // one set per image: the two image chains run on two streams
    float* desc1;  float* desc2;
    void* dense[2];   size_t dense_bytes;
    void* sink;    size_t sink_bytes;
    size_t total;
};

// The detector / descriptor chains of image 1 and image 2 are independent until the Sinkhorn kernel, and inside a chain
// the integral image needs only the image: with 4 streams image 2's chain and both integral builds run on side streams
// (fork and join by events, capturable into a CUDA graph), so that under-filled launches (top-k: one CTA per image),
// kernel tails and issue-bound / bandwidth-bound kernels overlap.
// Side streams and events belong to ONE caller stream: a small per-(device, caller stream) pool, so that calls on
// different caller streams (HostBatchMatcher's chunk streams, two caller threads) share nothing and overlap freely.
int g_match_binary = 1;                          // om_debug_match_binary: 0 = hard-binarised descriptors take the fp16 operand terms too
int g_match_streams = 4;                        // om_debug_match_streams: 1 = everything on the caller's stream, 2 = two chains, 4

struct SideSet {
    int device = -1;
    cudaStream_t owner = nullptr;               // the caller stream this set serves
    cudaStream_t side[3] = {};
    cudaEvent_t fork = nullptr, pre[2] = {}, join[3] = {};
    unsigned long long last_use = 0;
    std::mutex busy;                            // two host threads on the SAME caller stream: record / wait pairs must not interleave
};
constexpr int MAX_SIDE_SETS = 64;
SideSet g_sets[MAX_SIDE_SETS];
int g_num_sets = 0;
unsigned long long g_use_clock = 0;
std::mutex g_pool_mutex;                        // guards the table only, never held while work is enqueued

// the set serving (device, caller stream); created on first use, the least recently used one is re-assigned when the
// table is full (its streams are idle-or-ordered: a later caller simply queues behind whatever they still hold)
int side_set_for(int device, cudaStream_t owner, SideSet** out) {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    SideSet* lru = nullptr;
    for (int i = 0; i < g_num_sets; ++i) {
        SideSet& s = g_sets[i];
        if (s.device == device && s.owner == owner) { s.last_use = ++g_use_clock; *out = &s; return OM_OK; }
        if (s.device == device && (lru == nullptr || s.last_use < lru->last_use)) lru = &s;
    }
    SideSet* s = nullptr;
    if (g_num_sets < MAX_SIDE_SETS) {
        s = &g_sets[g_num_sets];
        for (int i = 0; i < 3; ++i) OM_CUDA(cudaStreamCreateWithFlags(&s->side[i], cudaStreamNonBlocking));
        OM_CUDA(cudaEventCreateWithFlags(&s->fork, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) OM_CUDA(cudaEventCreateWithFlags(&s->pre[i], cudaEventDisableTiming));
        for (int i = 0; i < 3; ++i) OM_CUDA(cudaEventCreateWithFlags(&s->join[i], cudaEventDisableTiming));
        s->device = device;
        ++g_num_sets;
    } else if (lru != nullptr) {
        s = lru;
    } else {
        return OM_ERR_LIMIT;
    }
    s->owner = owner;
    s->last_use = ++g_use_clock;
    *out = s;
    return OM_OK;
}

// Everything a stage launcher would reject is rejected HERE, before any work is forked onto side streams
int check_params(const om_match_params* p) {
    if (p == nullptr) return OM_ERR_NULL;
    if (p->flavour < OM_MATCH_SPARSE || p->flavour > OM_MATCH_MAPS) return OM_ERR_PARAM;
    if (p->B <= 0 || p->H <= 1 || p->W <= 1 || p->K <= 0) return OM_ERR_SHAPE;
    if ((long long)p->K > (long long)p->H * p->W) return OM_ERR_SHAPE;   // torch.topk raises too
    if ((long long)p->H * p->W >= (1ll << 31) || p->B > 65535 || p->K > OM_MAX_K) return OM_ERR_LIMIT;
    if (p->P != 256 && p->P != 512) return OM_ERR_PARAM;                 // descriptor/bad.py:385-388
    if (p->flavour != OM_MATCH_MAPS && (p->block_size < 1 || p->block_size % 2 == 0 || p->block_size / 2 > OM_MAX_BLOCK_HALF))
        return OM_ERR_PARAM;
    if (p->nms_radius < 0 || p->nms_radius > OM_MAX_NMS_RADIUS) return OM_ERR_PARAM;
    if (p->desc_mode < OM_DESC_RAW || p->desc_mode > OM_DESC_HARD) return OM_ERR_PARAM;
    if (p->flavour != OM_MATCH_DENSE && p->sampling_mode != OM_SAMPLE_NEAREST && p->sampling_mode != OM_SAMPLE_BILINEAR)
        return OM_ERR_PARAM;
    if (p->flavour == OM_MATCH_ANGLE && (p->patch_size < 1 || p->patch_size % 2 == 0 || p->patch_size > 31)) return OM_ERR_PARAM;
    if (p->iterations <= 0 || !(p->epsilon > 0.0f)) return OM_ERR_PARAM;  // matching/sinkhorn.py:66-69
    if (p->image_dtype != OM_IMAGE_F32 && p->image_dtype != OM_IMAGE_U8) return OM_ERR_PARAM;
    if (p->flavour == OM_MATCH_MAPS && p->image_dtype != OM_IMAGE_F32) return OM_ERR_PARAM;
    return OM_OK;
}

MatchWs plan(const om_match_params* p, void* base, bool ex = false) {
    MatchWs w{};
    char* c = (char*)base;
    size_t off = 0;
    w.detect_bytes = topk_workspace_bytes(p->B, p->H, p->W, p->K);
    w.detect[0] = c + off; off += align_up(w.detect_bytes);
    w.detect[1] = c + off; off += align_up(w.detect_bytes);
    const size_t db = align_up((size_t)p->B * p->K * p->P * sizeof(float));
    w.desc1 = (float*)(c + off); off += db;
    w.desc2 = (float*)(c + off); off += db;
    w.dense_bytes = p->flavour == OM_MATCH_DENSE
                        ? dense_bad_workspace_bytes(p->B, p->H, p->W)
                        : sparse_bad_workspace_bytes(p->B, p->H, p->W,
                                                     p->flavour == OM_MATCH_ANGLE || p->flavour == OM_MATCH_MAPS ? OM_THETA_MOMENTS
                                                                                                              : OM_THETA_NONE);
    w.dense[0] = c + off; off += align_up(w.dense_bytes);
    w.dense[1] = c + off; off += align_up(w.dense_bytes);
    w.sink_bytes = ex ? sinkhorn_ex_workspace_bytes(p->B, p->K, p->K, p->P) : sinkhorn_workspace_bytes(p->B, p->K, p->K, p->P);
    w.sink = c + off; off += align_up(w.sink_bytes);
    w.total = off;
    return w;
}

}  // namespace

namespace om {
int device_count_cached() {
    static int n = -1;
    if (n < 0) {
        int c = 0;
        n = cudaGetDeviceCount(&c) == cudaSuccess ? c : 0;
    }
    return n;
}
}  // namespace om

extern "C" void om_debug_match_streams(int n) { g_match_streams = n == 1 ? 1 : (n == 2 ? 2 : 4); }
extern "C" void om_debug_match_binary(int on) { g_match_binary = on ? 1 : 0; }

extern "C" size_t om_match_workspace_bytes(const om_match_params* p) {
    if (check_params(p) != OM_OK) return 0;
    return plan(p, nullptr).total;
}

extern "C" int om_match_pairs_f32(const om_match_params* p, const float* image1, const float* image2,
                                  const float* pair_table, const float* moment_kernels, float* kpts1, float* kpts2,
                                  float* probs, float* desc1, float* desc2, void* ws, size_t ws_bytes, void* stream) {
    if (p != nullptr && p->image_dtype != OM_IMAGE_F32) return OM_ERR_PARAM;
    return om_match_pairs(p, image1, image2, pair_table, moment_kernels, kpts1, kpts2, probs, desc1, desc2, ws, ws_bytes, stream);
}

namespace {
struct MatchMaps {                  // OM_MATCH_MAPS: the caller's detector outputs
    const float* scores[2];
    const float* orient[2];         // both null: non-oriented descriptors
};
int match_pairs_impl(const om_match_params* p, const void* image1, const void* image2, const float* pair_table,
                     const float* moment_kernels, float* kpts1, float* kpts2, float* probs, float* desc1, float* desc2,
                     const SinkhornEpilogue* epi, void* ws, size_t ws_bytes, void* stream, const MatchMaps* maps = nullptr);
}

extern "C" int om_match_pairs_from_maps_f32(const om_match_params* p, const float* image1, const float* image2,
                                            const float* scores1, const float* scores2, const float* orient1,
                                            const float* orient2, const float* pair_table, float* kpts1, float* kpts2,
                                            float* probs, float* desc1, float* desc2, void* ws, size_t ws_bytes, void* stream) {
    if (p == nullptr || probs == nullptr || scores1 == nullptr || scores2 == nullptr) return OM_ERR_NULL;
    if (p->flavour != OM_MATCH_MAPS) return OM_ERR_PARAM;
    if ((orient1 == nullptr) != (orient2 == nullptr)) return OM_ERR_NULL;
    const MatchMaps maps{{scores1, scores2}, {orient1, orient2}};
    return match_pairs_impl(p, image1, image2, pair_table, nullptr, kpts1, kpts2, probs, desc1, desc2, nullptr, ws, ws_bytes,
                            stream, &maps);
}

extern "C" int om_match_pairs(const om_match_params* p, const void* image1, const void* image2,
                              const float* pair_table, const float* moment_kernels, float* kpts1, float* kpts2,
                              float* probs, float* desc1, float* desc2, void* ws, size_t ws_bytes, void* stream) {
    if (probs == nullptr) return OM_ERR_NULL;
    return match_pairs_impl(p, image1, image2, pair_table, moment_kernels, kpts1, kpts2, probs, desc1, desc2, nullptr, ws, ws_bytes,
                            stream);
}

extern "C" size_t om_match_ex_workspace_bytes(const om_match_params* p) {
    if (check_params(p) != OM_OK) return 0;
    return plan(p, nullptr, true).total;
}

extern "C" int om_match_pairs_ex(const om_match_params* p, const void* image1, const void* image2, const float* pair_table,
                                 const float* moment_kernels, float* kpts1, float* kpts2, float* desc1, float* desc2,
                                 const om_sinkhorn_outputs* out, void* ws, size_t ws_bytes, void* stream) {
    if (out == nullptr) return OM_ERR_NULL;
    SinkhornEpilogue e = epilogue_from_outputs(out);
    e.kpts1 = kpts1;                                  // the matcher's own keypoints
    e.kpts2 = kpts2;
    if (!e.any() && out->probs == nullptr) return OM_ERR_NULL;
    return match_pairs_impl(p, image1, image2, pair_table, moment_kernels, kpts1, kpts2, out->probs, desc1, desc2, &e, ws, ws_bytes,
                            stream);
}

namespace {
int match_pairs_impl(const om_match_params* p, const void* image1, const void* image2, const float* pair_table,
                     const float* moment_kernels, float* kpts1, float* kpts2, float* probs, float* desc1, float* desc2,
                     const SinkhornEpilogue* epi, void* ws, size_t ws_bytes, void* stream, const MatchMaps* maps) {
    OM_ON_DEVICE_OF(image1);
    OM_TRY(check_params(p));
    if ((p->flavour == OM_MATCH_MAPS) != (maps != nullptr)) return OM_ERR_PARAM;
    if (image1 == nullptr || image2 == nullptr || pair_table == nullptr || kpts1 == nullptr || kpts2 == nullptr) return OM_ERR_NULL;
    if (p->flavour == OM_MATCH_ANGLE && moment_kernels == nullptr) return OM_ERR_NULL;
    const bool ex = epi != nullptr;
    if (ws == nullptr || ws_bytes < plan(p, nullptr, ex).total) return OM_ERR_WORKSPACE;
    const MatchWs w = plan(p, ws, ex);
    cudaStream_t st = (cudaStream_t)stream;
    float* d1 = desc1 ? desc1 : w.desc1;
    float* d2 = desc2 ? desc2 : w.desc2;
    const int u8 = p->image_dtype == OM_IMAGE_U8;
    const DetectCfg dc{p->B, p->H, p->W, p->block_size, p->nms_radius, p->border_margin, p->score_threshold, p->K, u8};
    const void* images[2] = {image1, image2};
    float* kp[2] = {kpts1, kpts2};
    float* ds[2] = {d1, d2};
    cudaStream_t chain[2] = {st, st};           // detector + descriptors of image 1 / image 2
    cudaStream_t pre[2] = {st, st};             // integral image of image 1 / image 2
    const int ns = g_match_streams;
    // hybrid Sinkhorn kernel ahead: its operand packing of image s runs on image s's own chain, right behind its descriptors
    const bool hy = sinkhorn_routes_to_hy(p->B, p->K, p->K, p->P, p->epsilon, p->unused_score, p->distance_l1, w.sink, w.sink_bytes) &&
                    p->iterations > 0 && (!ex || epi->any());
    if (hy && ex) {                                       // what sinkhorn_ex_launch would reject
        const SinkhornEpilogue& e = *epi;
        if (e.matches && (e.kpts1 == nullptr || e.kpts2 == nullptr || e.mk1 == nullptr || e.mk2 == nullptr || e.mscores == nullptr ||
                          e.mvalid == nullptr))
            return OM_ERR_NULL;
        if (e.matches && e.max_matches <= 0) return OM_ERR_SHAPE;
        if (e.filters && e.filter_valid == nullptr) return OM_ERR_NULL;
    }
    // the streaming kernels beyond 1024 keypoints pack their operands the same way (plain P output only)
    const bool xl = !hy && !ex && p->iterations > 0 && probs != nullptr &&
                    sinkhorn_routes_to_xl(p->B, p->K, p->K, p->P, p->epsilon, p->unused_score, p->distance_l1, w.sink, w.sink_bytes);
    // hard-binarised sparse descriptors are 0 / one value per row: the hybrid kernel then takes 8-bit operands (popcount GEMM)
    const int binary = hy && g_match_binary && p->desc_mode == OM_DESC_HARD && p->flavour != OM_MATCH_DENSE &&
                       sinkhorn_hy_binary_ok(p->K, p->K, p->P);
    if (hy) OM_TRY(sinkhorn_hy_prepare(p->B, p->K, p->K, p->P, w.sink, w.sink_bytes, st));   // before the fork: ordered ahead of both chains
    if (xl) OM_TRY(sinkhorn_xl_prepare(p->B, p->K, p->K, p->P, w.sink, w.sink_bytes, st));
    SideSet* set = nullptr;
    std::unique_lock<std::mutex> busy;
    int nside = 0;                              // side streams that were forked and must be joined, on every path
    if (ns > 1) {
        OM_TRY(side_set_for(dev_scope__.dev, st, &set));
        busy = std::unique_lock<std::mutex>(set->busy);
        chain[1] = set->side[0];
        if (ns == 4) { pre[0] = set->side[1]; pre[1] = set->side[2]; }
        OM_CUDA(cudaEventRecord(set->fork, st));                    // side work starts behind the caller's prior work
        nside = ns == 4 ? 3 : 1;
        for (int i = 0; i < nside; ++i) {
            const cudaError_t e = cudaStreamWaitEvent(set->side[i], set->fork, 0);
            if (e != cudaSuccess) { nside = i; break; }
        }
    }
    auto descriptors = [&](int s, cudaStream_t q, int phase) -> int {
        if (p->flavour == OM_MATCH_DENSE)
            return dense_bad_at_kpts_launch(images[s], u8, p->B, p->H, p->W, kp[s], p->K, pair_table, p->P, p->desc_mode,
                                            p->temperature, p->normalize, ds[s], w.dense[s], w.dense_bytes, q, phase);
        const int theta = p->flavour == OM_MATCH_ANGLE ? OM_THETA_MOMENTS
                          : (maps != nullptr && maps->orient[s] != nullptr ? OM_THETA_MAP : OM_THETA_NONE);
        return sparse_bad_launch(images[s], u8, p->B, p->H, p->W, kp[s], p->K, pair_table, p->P, p->desc_mode, p->temperature,
                                 p->normalize, p->sampling_mode, theta, maps != nullptr ? maps->orient[s] : nullptr, moment_kernels,
                                 p->patch_size, ds[s], w.dense[s], w.dense_bytes, q, phase);
    };
    // the forked part: whatever it returns, the side streams are joined back below (an unjoined fork would leave side work
    // running on a workspace the caller may free, and would invalidate a stream capture)
    auto forked = [&]() -> int {
        if (ns > 1 && nside != (ns == 4 ? 3 : 1)) return OM_ERR_CUDA_BASE + (int)cudaErrorUnknown;
        if (ns == 4) {
            for (int s = 0; s < 2; ++s) {
                OM_TRY(descriptors(s, pre[s], 1));                  // integral image only
                OM_CUDA(cudaEventRecord(set->pre[s], pre[s]));
            }
        }
        for (int s = 0; s < 2; ++s) {
            // keypoint scores are discarded by the matcher modules (`keypoints1, _ = ...`)
            if (maps != nullptr)
                OM_TRY(detect_scores_launch(maps->scores[s], dc, kp[s], nullptr, w.detect[s], w.detect_bytes, chain[s]));
            else
                OM_TRY(detect_launch(images[s], dc, nullptr, kp[s], nullptr, w.detect[s], w.detect_bytes, chain[s]));
            if (ns == 4) {
                OM_CUDA(cudaStreamWaitEvent(chain[s], set->pre[s], 0));
                OM_TRY(descriptors(s, chain[s], 2));
            } else {
                OM_TRY(descriptors(s, chain[s], 0));
            }
            if (hy) OM_TRY(sinkhorn_hy_pack(s, ds[s], p->B, p->K, p->K, p->P, w.sink, w.sink_bytes, chain[s], binary));
            if (xl) OM_TRY(sinkhorn_xl_pack(s, ds[s], p->B, p->K, p->K, p->P, w.sink, w.sink_bytes, chain[s]));
        }
        return OM_OK;
    };
    int rc = forked();
    for (int i = 0; i < nside; ++i) {                               // join: always, also on the error paths
        const cudaError_t e1 = cudaEventRecord(set->join[i], set->side[i]);
        const cudaError_t e2 = e1 == cudaSuccess ? cudaStreamWaitEvent(st, set->join[i], 0) : e1;
        if (e2 != cudaSuccess && rc == OM_OK) rc = OM_ERR_CUDA_BASE + (int)e2;
    }
    if (rc != OM_OK) return rc;
    if (hy)
        return sinkhorn_hy_run(d1, d2, p->B, p->K, p->K, p->P, p->iterations, p->epsilon, p->unused_score, probs, ex ? epi : nullptr,
                               w.sink, w.sink_bytes, st, binary);
    if (xl)
        return sinkhorn_xl_run(d1, d2, p->B, p->K, p->K, p->P, p->iterations, p->epsilon, p->unused_score, probs, w.sink, w.sink_bytes, st);
    if (ex)
        return sinkhorn_ex_launch(d1, d2, p->B, p->K, p->K, p->P, p->iterations, p->epsilon, p->unused_score, p->distance_l1,
                                  probs, *epi, w.sink, w.sink_bytes, st);
    return sinkhorn_launch(d1, d2, p->B, p->K, p->K, p->P, p->iterations, p->epsilon, p->unused_score, p->distance_l1,
                           probs, w.sink, w.sink_bytes, st);                     // needs both descriptor sets
}
}  // namespace
