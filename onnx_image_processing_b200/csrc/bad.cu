// BAD descriptor stage and orientation.
//
// Replaces (reference file:line):
//   descriptor/bad.py:436-576           SparseBAD.forward (non-oriented and oriented, nearest / bilinear)
//   descriptor/bad.py:62-110, 189-218   BADDescriptor dense map (float32 integral image, 4-tap order)
//   descriptor/bad.py:221-333           extract_descriptors_at_keypoints[_subpixel]
//   orientation/angle_estimation.py:123-172   AngleEstimator.forward
//
// Sparse BAD: the reference builds an 8-channel box-average bank with a 15x15 conv over the whole
// image and samples it with grid_sample.  Here one 64-thread group per keypoint stages the
// replicate-clamped patch around the keypoint in shared memory, turns it into an exact (fp64)
// integral image, and reads every box mean as 4 taps.  The sample coordinates go through the same
// fp32 normalise -> unnormalise -> clip -> round pipeline as grid_sample, with _rn intrinsics so
// that no mul+add is contracted.
//
// Dense BAD: the reference's integral image is float32 produced by two double-accumulated cumsums
// (each rounded once) -- reproduced as such, because box means then differ from exact maths by up
// to ~1 grey level and parity is against the reference, not the maths.
#include <math_constants.h>

#include <type_traits>

#include "common.cuh"
#include "pair_order.inc"

namespace om {

namespace {

constexpr int TPG = 64;          // threads per keypoint group
constexpr int MAXR = 7;          // largest box radius in the tables

// row pitch (elements) of an integral image with replicate padding `pad` and the zero column: a multiple of 4
// so that rows are 16-byte aligned and a keypoint's window is one TMA box
__host__ __device__ inline int ipitch(int W, int pad) { return (W + 2 * pad + 1 + 3) / 4 * 4; }
// pitch of the 16-bit integral: rows are whole 16-byte pieces of eight elements
__host__ __device__ inline int ipitch16(int W, int pad) { return (W + 2 * pad + 1 + 7) / 8 * 8; }

struct PairRow {
    float ox1, ox2, oy1, oy2, r, thr;
};

__device__ __forceinline__ PairRow load_pair(const float* table, int p) {
    const float* t = table + (size_t)p * 6;
    PairRow q;
    q.ox1 = __ldg(t + 0); q.ox2 = __ldg(t + 1); q.oy1 = __ldg(t + 2); q.oy2 = __ldg(t + 3);
    q.r = __ldg(t + 4); q.thr = __ldg(t + 5);
    return q;
}

// grid_sample(align_corners=True, padding_mode="border") coordinate pipeline on one axis:
// caller normalises with `pos * scale - 1` (bad.py:528-535), ATen unnormalises with
// (g + 1) * ((size-1)/2) and clips to [0, size-1].
__device__ __forceinline__ float sample_coord(float pos, float scale, float half_extent, float max_val) {
    const float g = __fsub_rn(__fmul_rn(pos, scale), 1.0f);
    const float u = __fmul_rn(__fadd_rn(g, 1.0f), half_extent);
    return fminf(fmaxf(u, 0.0f), max_val);
}

// slot -> pair index (pair_order.inc): the built-in orders for 256 / 512 pairs, the identity otherwise
__device__ __forceinline__ int pair_of_slot(int slot, int P) {
    return P == 256 ? (int)OM_PAIR_ORDER_256[slot] : (P == 512 ? (int)OM_PAIR_ORDER_512[slot] : slot);
}
// the same for the sparse window kernel on a 16-bit integral (another window pitch, two columns per bank word)
__device__ __forceinline__ int pair_of_slot_u16(int slot, int P) {
    return P == 256 ? (int)OM_PAIR_ORDER_256_U16[slot] : (P == 512 ? (int)OM_PAIR_ORDER_512_U16[slot] : slot);
}

__device__ __forceinline__ float finish_value(float diff, float thr, int mode, float temperature) {
    const float centered = __fsub_rn(diff, thr);                       // bad.py:212 / :559
    if (mode == OM_DESC_RAW) return centered;
    if (mode == OM_DESC_SOFT) {                                        // sigmoid(-centered * T)
        const float x = __fmul_rn(-centered, temperature);
        return 1.0f / (1.0f + expf(-x));
    }
    return centered <= 0.0f ? 1.0f : 0.0f;
}

struct SparseArgs {
    const void* image;         // float32, or uint8 when image_u8 (uint8 images are never flagged)
    int image_u8;
    int B, H, W;
    const float* kpts;
    int K;
    const float* table;
    int P;
    int mode;
    float temperature;
    int normalize;
    int bilinear;
    int theta_mode;
    const float* orientation;
    const float* moments;
    int patch_size;
    float* desc;
    float sy, sx;              // float32(2/(dim-1+1e-8)), bad.py:469-470
    const unsigned int* flags; // per image: 1 = has a pixel that is not an integer in [0,65535]; [B] = any image
    int pad;                   // replicate padding baked into the uint32 integral (window kernel)
};

// ---- sparse BAD, general kernel (any float image) ----------------------------------------------
// Used for the images the exact-uint32 window kernel below cannot take (flags[image] set).
// One 64-thread group (two warps, own named barrier) per keypoint, GROUPS keypoints per CTA.
//   1. patch -> exact integral image in shared memory (fp64: exact for integer AND float images).
//      Non-oriented: thread t owns patch column t, adds its S pixels top-down straight from global
//      memory (column prefix), then thread t owns row t (row prefix).  Oriented: the raw patch is
//      stored first because the 15x15 moments at the keypoint need the pixels themselves.
//   2. interior fast path (non-oriented, nearest, integer keypoint, patch fully inside the image --
//      ~88 % of keypoints at 480x640 with the default margin): nothing clamps, so the 8 integral taps
//      of every pair sit at keypoint-independent shared-memory offsets, precomputed once per CTA.
//      Everything else goes through the coordinate pipeline of grid_sample (general path).
//   3. box mean = float(exact sum) * float(1/area): the reference's bank weights are exactly
//      float(1/area) (bad.py:426-434), so this is its arithmetic with an exact accumulator.
// HS: patch half size.  Nearest / bilinear sample centres are at most 16 px from the rounded
// keypoint (15 offset + two roundings), + radius 7 -> HS = 23; rotated offsets reach 22 + 7 -> 30.
// correctly rounded s / area given y = float(1/area) (Markstein: one residual correction of the product; checked
// against exact rational arithmetic for every box area) -- replaces the ~15-instruction branchy IEEE division
__device__ __forceinline__ float div_area(float s, float area, float y) {
    const float q0 = __fmul_rn(s, y);
    const float r = __fmaf_rn(-area, q0, s);
    return __fmaf_rn(r, y, q0);
}

__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(TPG) : "memory"); }

// sum of `v` over the TPG threads of one keypoint group; red is per-CTA scratch [groups][2]
__device__ __forceinline__ float group_sum(float v, float* red, int g, int t) {
    v = warp_sum(v);
    group_bar(g);
    if ((t & 31) == 0) red[g * 2 + (t >> 5)] = v;
    group_bar(g);
    return red[g * 2] + red[g * 2 + 1];
}

// one keypoint, one 64-thread group (see sparse_bad_kernel below)
template <int HS, bool ORIENTED, bool BILINEAR>
__device__ __forceinline__ void sparse_bad_group(const SparseArgs& a, long long kidx, double* D, float* red, float* sTheta,
                                                 const uint4* sTap, const float2* sThr, int g, int t) {
    constexpr int S = 2 * HS + 1;       // patch side
    constexpr int PD = (S + 1) | 1;     // odd pitch (in doubles) of the (S+1)x(S+1) integral
    constexpr bool FAST = !ORIENTED && !BILINEAR;
    const int z = (int)(kidx / a.K);
    if (a.flags != nullptr && a.flags[z] == 0u) return;   // integer-valued image: sparse_win_kernel did this keypoint
    const int H = a.H, W = a.W;
    const float* img = reinterpret_cast<const float*>(a.image) + (size_t)z * H * W;   // flagged images are float32
    float* out = a.desc + (size_t)kidx * a.P;

    const float ky = a.kpts[kidx * 2 + 0], kx = a.kpts[kidx * 2 + 1];
    if (!(ky >= 0.0f)) {                                                    // bad.py:461, :570 -> the row is all zeros
        for (int p = t; p < a.P; p += TPG) out[p] = 0.0f;
        return;
    }
    const float yc = fminf(fmaxf(ky, 0.0f), (float)(H - 1));               // bad.py:464-465
    const float xc = fminf(fmaxf(kx, 0.0f), (float)(W - 1));
    const int iy0 = (int)nearbyintf(yc), ix0 = (int)nearbyintf(xc);
    const int py0 = iy0 - HS, px0 = ix0 - HS;                               // global coords of patch (0,0)
    const bool inside = py0 >= 0 && px0 >= 0 && py0 + S <= H && px0 + S <= W;

    float ct = 1.0f, st = 0.0f;
    if (!ORIENTED) {
        // column prefix straight from global memory; row 0 / column 0 of the integral are zero
        if (t < S) {
            D[t + 1] = 0.0;
            double acc = 0.0;
            double* col = D + PD + t + 1;
            if (inside) {
                const float* src = img + (size_t)py0 * W + px0 + t;
#pragma unroll
                for (int dy = 0; dy < S; ++dy) { acc += (double)__ldg(src + (size_t)dy * W); col[dy * PD] = acc; }
            } else {
                // replicate-clamped patch (bad.py:474-478 pads the image, grid_sample clamps the centre)
                const float* src = img + clampi(px0 + t, 0, W - 1);
#pragma unroll 8
                for (int dy = 0; dy < S; ++dy) {
                    acc += (double)__ldg(src + (size_t)clampi(py0 + dy, 0, H - 1) * W);
                    col[dy * PD] = acc;
                }
            }
        } else if (t == S) {
#pragma unroll 8
            for (int dy = 0; dy <= S; ++dy) D[dy * PD] = 0.0;
        }
        group_bar(g);
    } else {
        static_assert(S + 1 <= TPG, "one thread per integral column");
        if (t <= S) {
            double* col = D + t;
            col[0] = 0.0;
            if (t == 0) {
#pragma unroll 4
                for (int dy = 1; dy <= S; ++dy) col[dy * PD] = 0.0;
            } else {
                const float* src = img + clampi(px0 + t - 1, 0, W - 1);
#pragma unroll 8
                for (int dy = 1; dy <= S; ++dy)
                    col[dy * PD] = (double)__ldg(src + (size_t)clampi(py0 + dy - 1, 0, H - 1) * W);
            }
        }
        group_bar(g);
        // orientation of this keypoint: nearest sample of the orientation map (bad.py:490-499)
        const int ny = (int)nearbyintf(sample_coord(yc, a.sy, (float)(H - 1) * 0.5f, (float)(H - 1)));
        const int nx = (int)nearbyintf(sample_coord(xc, a.sx, (float)(W - 1) * 0.5f, (float)(W - 1)));
        float theta = 0.0f;
        if (a.theta_mode == OM_THETA_MAP) {
            theta = __ldg(a.orientation + (size_t)z * H * W + (size_t)ny * W + nx);
        } else {
            // angle_estimation.py:161-170 at (ny,nx) only: zero-padded cross-correlation, then atan2
            const int ps = a.patch_size, half = ps / 2;
            float m10 = 0.0f, m01 = 0.0f;
            for (int tap = t; tap < ps * ps; tap += TPG) {
                const int j = tap / ps, i = tap % ps;
                const int gy = ny + j - half, gx = nx + i - half;
                if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                    const int dy = clampi(gy - py0 + 1, 1, S), dx = clampi(gx - px0 + 1, 1, S);
                    const float v = (float)D[dy * PD + dx];
                    m10 = fmaf(__ldg(a.moments + tap), v, m10);
                    m01 = fmaf(__ldg(a.moments + ps * ps + tap), v, m01);
                }
            }
            m10 = group_sum(m10, red, g, t);
            m01 = group_sum(m01, red, g, t);
            theta = atan2f(m01, m10);
            if (t == 0) sTheta[g] = theta;      // one value for the whole group (sums are identical anyway)
            group_bar(g);
            theta = sTheta[g];
        }
        ct = cosf(theta);                                                   // bad.py:501-502
        st = sinf(theta);
        group_bar(g);
        if (t < S) {                                                        // column prefix
            double acc = 0.0;
            double* p = D + PD + (t + 1);
#pragma unroll 8
            for (int y = 0; y < S; ++y, p += PD) { acc += *p; *p = acc; }
        }
        group_bar(g);
    }
    if (t < S) {                                                            // row prefix
        double acc = 0.0;
        double* p = D + (t + 1) * PD + 1;
#pragma unroll
        for (int x = 0; x < S; ++x) { acc += p[x]; p[x] = acc; }
    }
    group_bar(g);

    constexpr int MAXPP = 8;   // pairs per thread: P <= 512
    float d[MAXPP];
    float ss = 0.0f;
    const bool fast = FAST && inside && ky == (float)iy0 && kx == (float)ix0;
    if (fast) {
#pragma unroll
        for (int q = 0; q < MAXPP; ++q) {
            const int p = t + q * TPG;
            d[q] = 0.0f;
            if (p < a.P) {
                const uint4 tp = sTap[p];
                const float2 tb = sThr[p];
                const double s1 = (D[tp.x & 0xFFFFu] - D[tp.x >> 16]) - (D[tp.y & 0xFFFFu] - D[tp.y >> 16]);
                const double s2 = (D[tp.z & 0xFFFFu] - D[tp.z >> 16]) - (D[tp.w & 0xFFFFu] - D[tp.w >> 16]);
                const float diff = __fsub_rn(__fmul_rn((float)s1, tb.y), __fmul_rn((float)s2, tb.y));   // bad.py:557
                d[q] = finish_value(diff, tb.x, a.mode, a.temperature);
                ss = fmaf(d[q], d[q], ss);
            }
        }
    } else {
        const float hy = (float)(H - 1) * 0.5f, hx = (float)(W - 1) * 0.5f;
        const float my = (float)(H - 1), mx = (float)(W - 1);
        auto box_mean = [&](int cy, int cx, int r, float inv_area) -> float {
            // cy,cx: integer sample centre in image coords; box of radius r on the replicate-padded image
            const int pyc = clampi(cy - py0, r, S - 1 - r), pxc = clampi(cx - px0, r, S - 1 - r);
            const int y0 = pyc - r, y1 = pyc + r + 1, x0 = pxc - r, x1 = pxc + r + 1;
            const double s = (D[y1 * PD + x1] - D[y0 * PD + x1]) - (D[y1 * PD + x0] - D[y0 * PD + x0]);
            return __fmul_rn((float)s, inv_area);
        };
        auto sample = [&](float oy, float ox, int r, float inv_area) -> float {
            float py, px;
            if (ORIENTED) {                                                     // bad.py:504-517
                const float dy = __fadd_rn(__fmul_rn(ox, st), __fmul_rn(oy, ct));
                const float dx = __fsub_rn(__fmul_rn(ox, ct), __fmul_rn(oy, st));
                py = __fadd_rn(yc, dy);
                px = __fadd_rn(xc, dx);
            } else {                                                            // bad.py:518-525
                py = __fadd_rn(yc, oy);
                px = __fadd_rn(xc, ox);
            }
            const float uy = sample_coord(py, a.sy, hy, my), ux = sample_coord(px, a.sx, hx, mx);
            if (!BILINEAR) return box_mean((int)nearbyintf(uy), (int)nearbyintf(ux), r, inv_area);   // half-to-even
            const float fy = floorf(uy), fx = floorf(ux);
            const float w = ux - fx, e = 1.0f - w, s = uy - fy, n = 1.0f - s;   // ATen bilinear weights
            const int y_n = (int)fy, x_w = (int)fx;
            const int y_s = min(y_n + 1, H - 1), x_e = min(x_w + 1, W - 1);     // weight is 0 where this clamps
            const float nw = box_mean(y_n, x_w, r, inv_area), ne = box_mean(y_n, x_e, r, inv_area);
            const float sw = box_mean(y_s, x_w, r, inv_area), se = box_mean(y_s, x_e, r, inv_area);
            return nw * (n * e) + ne * (n * w) + sw * (s * e) + se * (s * w);
        };
#pragma unroll
        for (int q = 0; q < MAXPP; ++q) {
            const int p = t + q * TPG;
            d[q] = 0.0f;
            if (p < a.P) {
                const PairRow row = load_pair(a.table, p);
                const int r = (int)row.r;
                const float side = (float)(2 * r + 1);
                const float inv_area = __fdiv_rn(1.0f, side * side);
                const float diff = __fsub_rn(sample(row.oy1, row.ox1, r, inv_area), sample(row.oy2, row.ox2, r, inv_area));
                d[q] = finish_value(diff, row.thr, a.mode, a.temperature);
                ss = fmaf(d[q], d[q], ss);
            }
        }
    }
    float inv = 1.0f;
    if (a.normalize) {                                                      // F.normalize, bad.py:573-574
        const float nrm = sqrtf(group_sum(ss, red, g, t));
        inv = 1.0f / fmaxf(nrm, 1e-12f);
    }
#pragma unroll
    for (int q = 0; q < MAXPP; ++q) {
        const int p = t + q * TPG;
        if (p < a.P) out[p] = d[q] * inv;
    }
}

template <int HS, int GROUPS, bool ORIENTED, bool BILINEAR>
__global__ void __launch_bounds__(GROUPS * TPG) sparse_bad_kernel(SparseArgs a) {
    if (a.flags != nullptr && a.flags[a.B] == 0u) return;   // whole batch integer-valued: nothing to do here
    constexpr int S = 2 * HS + 1;       // patch side
    constexpr int PD = (S + 1) | 1;     // odd pitch (in doubles) of the (S+1)x(S+1) integral
    constexpr bool FAST = !ORIENTED && !BILINEAR;
    extern __shared__ __align__(16) double sD[];
    __shared__ float red[GROUPS * 2];
    __shared__ float sTheta[GROUPS];
    // fast-path tables, one entry per pair: 8 integral offsets (uint16) and {threshold, 1/area}
    uint4* sTap = reinterpret_cast<uint4*>(sD + (size_t)GROUPS * (S + 1) * PD);
    float2* sThr = reinterpret_cast<float2*>(sTap + (FAST ? a.P : 0));

    const int g = threadIdx.x / TPG, t = threadIdx.x % TPG;
    if (FAST) {
        for (int p = threadIdx.x; p < a.P; p += GROUPS * TPG) {
            const PairRow row = load_pair(a.table, p);
            const int r = (int)row.r;
            const int cy1 = HS + (int)row.oy1, cx1 = HS + (int)row.ox1, cy2 = HS + (int)row.oy2, cx2 = HS + (int)row.ox2;
            auto off = [&](int y, int x) -> unsigned { return (unsigned)(y * PD + x); };
            uint4 tp;
            tp.x = off(cy1 + r + 1, cx1 + r + 1) | (off(cy1 - r, cx1 + r + 1) << 16);
            tp.y = off(cy1 + r + 1, cx1 - r) | (off(cy1 - r, cx1 - r) << 16);
            tp.z = off(cy2 + r + 1, cx2 + r + 1) | (off(cy2 - r, cx2 + r + 1) << 16);
            tp.w = off(cy2 + r + 1, cx2 - r) | (off(cy2 - r, cx2 - r) << 16);
            sTap[p] = tp;
            const float side = (float)(2 * r + 1);
            sThr[p] = make_float2(row.thr, __fdiv_rn(1.0f, side * side));
        }
        __syncthreads();
    }

    // persistent loop: the grid is small, so an all-integer batch (nothing to do here) costs ~1 us
    const long long total = (long long)a.B * a.K;
    const long long nblk = (total + GROUPS - 1) / GROUPS;
    double* D = sD + (size_t)g * (S + 1) * PD;
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const long long kidx = blk * GROUPS + g;
        if (kidx < total) sparse_bad_group<HS, ORIENTED, BILINEAR>(a, kidx, D, red, sTheta, sTap, sThr, g, t);
        group_bar(g);                   // the group's shared memory is reused by its next keypoint
    }
}

template <int HS, int GROUPS, bool ORIENTED, bool BILINEAR>
int launch_sparse(const SparseArgs& a, cudaStream_t st) {
    constexpr int S = 2 * HS + 1;
    constexpr int PD = (S + 1) | 1;
    const size_t smem = (size_t)GROUPS * (S + 1) * PD * sizeof(double) +
                        ((!ORIENTED && !BILINEAR) ? (size_t)a.P * (sizeof(uint4) + sizeof(float2)) : 0);
    auto kernel = sparse_bad_kernel<HS, GROUPS, ORIENTED, BILINEAR>;
    OM_TRY(set_smem(kernel, smem));
    const long long total = (long long)a.B * a.K;
    const long long nblk = (total + GROUPS - 1) / GROUPS;
    const unsigned grid = (unsigned)(nblk < 148 * 2 ? nblk : 148 * 2);
    kernel<<<grid, GROUPS * TPG, smem, st>>>(a);
    OM_AFTER_LAUNCH();
    return OM_OK;
}

// ---- sparse BAD, window kernel (the fast path) ------------------------------------------------------
// The exact uint32 integral of the replicate-padded image exists in global memory (build_prefix<true>),
// so a keypoint needs no arithmetic to get its patch: one TMA box load brings the (S+1)x(S+1) window of
// the integral around the keypoint into shared memory (padding is baked into the integral: the window
// is always in bounds), and every box sum is 4 taps in wrap-around uint32 arithmetic.
//   * interior fast path (non-oriented, nearest, integer keypoint at least 15 px from the border, i.e.
//     no sample centre clamps): the 8 taps of every pair sit at keypoint-independent window offsets,
//     precomputed once per CTA.
//   * general path: the fp32 normalise -> unnormalise -> clip -> round pipeline of grid_sample.
// box mean = float(exact sum) * float(1/area): the reference's bank weights are exactly float(1/area)
// (bad.py:426-434), so this is its arithmetic with an exact accumulator.
// Images whose flag is set (a pixel that is not an integer in [0,65535]) are skipped here and handled by
// sparse_bad_kernel (fp64 integral built per keypoint).
// Window geometry: nearest / bilinear sample centres are at most 16 px from the rounded keypoint (15
// offset + two roundings), + radius 7 -> HS = 23, window 48x48; rotated offsets reach 22 + 7 -> HS = 30,
// window 62 rows x 64 columns (box width must be a multiple of 16 bytes).
// geometry shared by the kernel and its launcher
// WT: element type of the integral: uint32, or uint16 (the integral modulo 2^16: every box sum is below 2^16 -- at most 15 x 15
// pixels of at most 255 (the integer-image test bounds the pixels) -- so wrap-around differences are exact in 16 bits too,
// and the window a keypoint fetches from L2 is half the bytes)
template <int HS, typename WT = unsigned int>
struct WinGeom {
    static constexpr int S = 2 * HS + 1;
    static constexpr int WR = S + 1;                    // window rows
    static constexpr int EPW = 16 / (int)sizeof(WT);    // elements per 16 bytes
    static constexpr int WP = (S + 1 + 2 * (EPW - 1)) / EPW * EPW;   // TMA box width == shared-memory pitch: the box must start
                                                        // on a 16-byte boundary in global memory, so up to EPW - 1 columns early
    static constexpr int GSTRIDE = (WR * WP * (int)sizeof(WT) / 4 + 31) / 32 * 32;   // 32-bit words per window buffer (128-byte aligned)
};

// The eight window byte offsets of a pair's taps (two boxes x four corners, bad.py:98 order), 16 bits each, in ONE 16-byte
// word per slot: a warp reads it as one conflict-free LDS.128 (two uint4 per slot at a 32-byte lane stride cost 16
// wavefronts instead of 4; the tables were 40 % of the descriptor kernels' shared-memory wavefronts).
__device__ __forceinline__ uint4 pack_taps(const unsigned (&o)[8]) {
    return make_uint4(o[0] | (o[1] << 16), o[2] | (o[3] << 16), o[4] | (o[5] << 16), o[6] | (o[7] << 16));
}
__device__ __forceinline__ unsigned tap_lo(unsigned v) { return v & 0xFFFFu; }
__device__ __forceinline__ unsigned tap_hi(unsigned v) { return v >> 16; }

// keypoint -> does it need a window here?  (valid, and its image is integer-valued)
__device__ __forceinline__ bool win_needed(const SparseArgs& a, long long kidx, float& ky, float& kx, int& z) {
    z = (int)(kidx / a.K);
    ky = a.kpts[kidx * 2 + 0];
    kx = a.kpts[kidx * 2 + 1];
    return ky >= 0.0f && a.flags[z] == 0u;
}

// per-keypoint context of the general (border / oriented / bilinear) pair evaluation
template <typename WT>
struct SparseKpCtx {
    const WT* Wn;                // window column 0
    float yc, xc, ct, st;        // clamped keypoint, cos / sin of its orientation
    int iy0, ix0;                // rounded keypoint
    float sy, sx;                // grid_sample scales
    int H, W;
};

// One pair at one keypoint through the full sampling pipeline (bad.py:504-557).
// `row`, `r`, `inv_area`: the pair's table entry and what follows from it (1 / box area), loaded by the caller (from the
// global table on the rare border path of the fast kernels, from the CTA's shared-memory copy in the oriented / bilinear
// kernels, where every pair of every keypoint comes through here).
template <int HS, bool ORIENTED, bool BILINEAR, typename WT>
__device__ __forceinline__ float sparse_pair_general_impl(const SparseKpCtx<WT>& c, const PairRow& row, int r, float inv_area, int mode,
                                                          float temperature) {
    using G = WinGeom<HS, WT>;
    constexpr int S = G::S, WP = G::WP;
    const int H = c.H, W = c.W;
    const float hy = (float)(H - 1) * 0.5f, hx = (float)(W - 1) * 0.5f;
    const float my = (float)(H - 1), mx = (float)(W - 1);
    const WT* Wn = c.Wn;
    auto box_mean = [&](int cy, int cx, int r, float inv_area) -> float {
        // cy,cx: integer sample centre in image coords (already clipped to the image by grid_sample's
        // border mode); the box may reach into the replicate padding, which the integral contains
        const int wy = clampi(cy - c.iy0 + HS, r, S - 1 - r), wx = clampi(cx - c.ix0 + HS, r, S - 1 - r);
        const int y0 = wy - r, y1 = wy + r + 1, x0 = wx - r, x1 = wx + r + 1;
        unsigned int s = ((unsigned int)Wn[y1 * WP + x1] - (unsigned int)Wn[y0 * WP + x1]) -
                         ((unsigned int)Wn[y1 * WP + x0] - (unsigned int)Wn[y0 * WP + x0]);
        if (sizeof(WT) == 2) s &= 0xFFFFu;                                  // the integral modulo 2^16
        return __fmul_rn((float)s, inv_area);
    };
    auto sample = [&](float oy, float ox, int r, float inv_area) -> float {
        float py, px;
        if (ORIENTED) {                                                     // bad.py:504-517
            const float dy = __fadd_rn(__fmul_rn(ox, c.st), __fmul_rn(oy, c.ct));
            const float dx = __fsub_rn(__fmul_rn(ox, c.ct), __fmul_rn(oy, c.st));
            py = __fadd_rn(c.yc, dy);
            px = __fadd_rn(c.xc, dx);
        } else {                                                            // bad.py:518-525
            py = __fadd_rn(c.yc, oy);
            px = __fadd_rn(c.xc, ox);
        }
        const float uy = sample_coord(py, c.sy, hy, my), ux = sample_coord(px, c.sx, hx, mx);
        if (!BILINEAR) return box_mean((int)nearbyintf(uy), (int)nearbyintf(ux), r, inv_area);   // half-to-even
        const float fy = floorf(uy), fx = floorf(ux);
        const float w = ux - fx, e = 1.0f - w, s = uy - fy, n = 1.0f - s;   // ATen bilinear weights
        const int y_n = (int)fy, x_w = (int)fx;
        const int y_s = min(y_n + 1, H - 1), x_e = min(x_w + 1, W - 1);     // weight is 0 where this clamps
        const float nw = box_mean(y_n, x_w, r, inv_area), ne = box_mean(y_n, x_e, r, inv_area);
        const float sw = box_mean(y_s, x_w, r, inv_area), se = box_mean(y_s, x_e, r, inv_area);
        return nw * (n * e) + ne * (n * w) + sw * (s * e) + se * (s * w);
    };
    const float diff = __fsub_rn(sample(row.oy1, row.ox1, r, inv_area), sample(row.oy2, row.ox2, r, inv_area));
    return finish_value(diff, row.thr, mode, temperature);
}
// Out-of-line form for the kernels that have the tap-table fast path (nearest, not oriented): there this is the rare
// border path, and inlined into the unrolled pair loop next to the fast path it made the kernel several hundred KB of
// code (stalls on instruction fetch).  The oriented / bilinear kernels have only this path and inline it, so that the
// pairs of a thread overlap.
template <int HS, bool ORIENTED, bool BILINEAR, typename WT>
__device__ __noinline__ float sparse_pair_general(const SparseKpCtx<WT>& c, const float* table, int p, int mode, float temperature) {
    const PairRow row = load_pair(table, p);
    const int r = (int)row.r;
    const float side = (float)(2 * r + 1);
    return sparse_pair_general_impl<HS, ORIENTED, BILINEAR, WT>(c, row, r, __fdiv_rn(1.0f, side * side), mode, temperature);
}

// one keypoint whose window `win` is in flight / has landed on mbarrier `bar` (phase `parity`).  NPP = pairs per
// thread (ceil(P / 64)): the pair loops are unrolled over exactly NPP slots.
template <int HS, bool ORIENTED, bool BILINEAR, int NPP, typename WT>
__device__ __forceinline__ void sparse_win_group(const SparseArgs& a, long long kidx, int z, float ky, float kx,
                                                 const WT* win, uint32_t bar, uint32_t parity, float* red,
                                                 float* sTheta, const uint4* sTap, const float2* sThr,
                                                 const unsigned short* sIdx, const float4* sOff, const float4* sPar, int g, int t) {
    constexpr bool FAST = !ORIENTED && !BILINEAR;
    const int H = a.H, W = a.W;
    float* out = a.desc + (size_t)kidx * a.P;
    const float yc = fminf(fmaxf(ky, 0.0f), (float)(H - 1));               // bad.py:464-465
    const float xc = fminf(fmaxf(kx, 0.0f), (float)(W - 1));
    const int iy0 = (int)nearbyintf(yc), ix0 = (int)nearbyintf(xc);
    const int wx0 = ix0 - HS + a.pad;            // integral column of window column 0 (>= 0)
    const WT* Wn = win + (wx0 & (WinGeom<HS, WT>::EPW - 1));    // the box starts on a 16-byte boundary, up to 3 (7) columns early

    float ct = 1.0f, st = 0.0f;
    if (ORIENTED) {
        // orientation of this keypoint: nearest sample of the orientation map (bad.py:490-499)
        const int ny = (int)nearbyintf(sample_coord(yc, a.sy, (float)(H - 1) * 0.5f, (float)(H - 1)));
        const int nx = (int)nearbyintf(sample_coord(xc, a.sx, (float)(W - 1) * 0.5f, (float)(W - 1)));
        float theta;
        if (a.theta_mode == OM_THETA_MAP) {
            theta = __ldg(a.orientation + (size_t)z * H * W + (size_t)ny * W + nx);
        } else {
            // angle_estimation.py:161-170 at (ny,nx) only: zero-padded cross-correlation, then atan2
            const size_t ibase = (size_t)z * H * W;
            const float* imgf = reinterpret_cast<const float*>(a.image);
            const unsigned char* imgb = reinterpret_cast<const unsigned char*>(a.image);
            const int ps = a.patch_size, half = ps / 2;
            float m10 = 0.0f, m01 = 0.0f;
            for (int tap = t; tap < ps * ps; tap += TPG) {
                const int j = tap / ps, i = tap - j * ps;
                const int gy = ny + j - half, gx = nx + i - half;
                if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                    const size_t idx = ibase + (size_t)gy * W + gx;
                    const float v = a.image_u8 ? (float)__ldg(imgb + idx) : __ldg(imgf + idx);
                    m10 = fmaf(__ldg(a.moments + tap), v, m10);
                    m01 = fmaf(__ldg(a.moments + ps * ps + tap), v, m01);
                }
            }
            m10 = group_sum(m10, red, g, t);
            m01 = group_sum(m01, red, g, t);
            theta = atan2f(m01, m10);
            if (t == 0) sTheta[g] = theta;
            group_bar(g);
            theta = sTheta[g];
        }
        ct = cosf(theta);                                                   // bad.py:501-502
        st = sinf(theta);
    }
    mbar_wait(bar, parity);                      // window has landed

    float d[NPP];
    float ss = 0.0f;
    const bool fast = FAST && ky == (float)iy0 && kx == (float)ix0 && iy0 >= 15 && iy0 + 14 <= H - 1 && ix0 >= 15 &&
                      ix0 + 14 <= W - 1;
    const char* wbytes = reinterpret_cast<const char*>(Wn);
    auto ldw = [&](unsigned int byte_off) -> unsigned int { return (unsigned int)*reinterpret_cast<const WT*>(wbytes + byte_off); };
    constexpr unsigned int WMASK = sizeof(WT) == 2 ? 0xFFFFu : 0xFFFFFFFFu;     // 16-bit integral: sums modulo 2^16
    if (fast) {
#pragma unroll
        for (int q = 0; q < NPP; ++q) {
            const int p = t + q * TPG;
            d[q] = 0.0f;
            if (p < a.P) {
                const uint4 tp = sTap[p];
                const float2 tb = sThr[p];
                const unsigned int s1 = ((ldw(tap_lo(tp.x)) - ldw(tap_hi(tp.x))) - (ldw(tap_lo(tp.y)) - ldw(tap_hi(tp.y)))) & WMASK;
                const unsigned int s2 = ((ldw(tap_lo(tp.z)) - ldw(tap_hi(tp.z))) - (ldw(tap_lo(tp.w)) - ldw(tap_hi(tp.w)))) & WMASK;
                const float diff = __fsub_rn(__fmul_rn((float)s1, tb.y), __fmul_rn((float)s2, tb.y));   // bad.py:557
                d[q] = finish_value(diff, tb.x, a.mode, a.temperature);
                ss = fmaf(d[q], d[q], ss);
            }
        }
    } else {
        SparseKpCtx<WT> c;
        c.Wn = Wn; c.yc = yc; c.xc = xc; c.ct = ct; c.st = st; c.iy0 = iy0; c.ix0 = ix0; c.sy = a.sy; c.sx = a.sx; c.H = H; c.W = W;
#pragma unroll
        for (int q = 0; q < NPP; ++q) {
            const int p = t + q * TPG;
            d[q] = 0.0f;
            if (p < a.P) {
                if (FAST) {
                    d[q] = sparse_pair_general<HS, ORIENTED, BILINEAR, WT>(c, a.table, p, a.mode, a.temperature);
                } else {
                    const float4 o = sOff[p], pr = sPar[p];                 // {oy1, ox1, oy2, ox2}, {threshold, 1/area, r, -}
                    PairRow row;
                    row.oy1 = o.x; row.ox1 = o.y; row.oy2 = o.z; row.ox2 = o.w; row.thr = pr.x; row.r = 0.0f;
                    d[q] = sparse_pair_general_impl<HS, ORIENTED, BILINEAR, WT>(c, row, __float_as_int(pr.z), pr.y, a.mode, a.temperature);
                }
                ss = fmaf(d[q], d[q], ss);
            }
        }
    }
    float inv = 1.0f;
    if (a.normalize) {                                                      // F.normalize, bad.py:573-574
        const float nrm = sqrtf(group_sum(ss, red, g, t));
        inv = 1.0f / fmaxf(nrm, 1e-12f);
    }
#pragma unroll
    for (int q = 0; q < NPP; ++q) {
        const int p = t + q * TPG;
        if (p < a.P) out[fast ? (int)sIdx[p] : p] = d[q] * inv;     // fast path: slot p holds pair sIdx[p]
    }
}

// GROUPS keypoint groups per CTA; every group walks its keypoints (grid-stride).  NBUF = 2: two-deep window
// pipeline, the TMA box of keypoint n+1 is in flight while keypoint n is evaluated; NBUF = 1: one window per group
// and twice the resident groups instead (measured faster on B200: the kernel is bound by shared-memory wavefronts
// of the random taps, not by the TMA latency, so more warps beat prefetching).
template <int HS, int GROUPS, bool ORIENTED, bool BILINEAR, int NBUF, int NPP, typename WT>
__global__ void __launch_bounds__(GROUPS * TPG, 4) sparse_win_kernel(const __grid_constant__ CUtensorMap tmap, SparseArgs a) {
    using G = WinGeom<HS, WT>;
    constexpr bool FAST = !ORIENTED && !BILINEAR;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ float red[GROUPS * 2];
    __shared__ float sTheta[GROUPS];
    __shared__ __align__(8) unsigned long long bars[GROUPS * 2];
    unsigned int* sWin = reinterpret_cast<unsigned int*>(smem_raw);             // GROUPS x NBUF x GSTRIDE
    uint4* sTap = reinterpret_cast<uint4*>(sWin + GROUPS * NBUF * G::GSTRIDE);  // fast path: 2 x 4 window byte offsets per pair (16 bits each)
    float2* sThr = reinterpret_cast<float2*>(sTap + (FAST ? a.P : 0));          //            {threshold, 1/area}
    unsigned short* sIdx = reinterpret_cast<unsigned short*>(sThr + (FAST ? a.P : 0));   //     slot -> pair (pair_order.inc)
    // oriented / bilinear kernels: the pair table itself (every keypoint evaluates every pair through the general path)
    float4* sOff = reinterpret_cast<float4*>(sWin + GROUPS * NBUF * G::GSTRIDE);           // {oy1, ox1, oy2, ox2}
    float4* sPar = sOff + (FAST ? 0 : a.P);                                                // {threshold, 1/area, r, -}

    const int g = threadIdx.x / TPG, t = threadIdx.x % TPG;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < GROUPS * 2; ++i) mbar_init(smem_u32(&bars[i]), 1);
        mbar_fence_init();
    }
    if (FAST) {
        for (int p = threadIdx.x; p < a.P; p += GROUPS * TPG) {
            // slot p takes pair pair_of_slot(p): the 32 lanes of a tap read are 32 consecutive slots, ordered for few bank
            // conflicts (the window geometry differs from the dense kernel's by a constant offset: same order)
            const int pair = sizeof(WT) == 2 ? pair_of_slot_u16(p, a.P) : pair_of_slot(p, a.P);
            sIdx[p] = (unsigned short)pair;
            const PairRow row = load_pair(a.table, pair);
            const int r = (int)row.r;
            const int cy1 = HS + (int)row.oy1, cx1 = HS + (int)row.ox1, cy2 = HS + (int)row.oy2, cx2 = HS + (int)row.ox2;
            auto off = [&](int y, int x) -> unsigned { return (unsigned)(y * G::WP + x) * (unsigned)sizeof(WT); };
            static_assert(G::WR * G::WP * sizeof(WT) < 65536, "window byte offsets must fit 16 bits");
            const unsigned o[8] = {off(cy1 + r + 1, cx1 + r + 1), off(cy1 - r, cx1 + r + 1), off(cy1 + r + 1, cx1 - r), off(cy1 - r, cx1 - r),
                                   off(cy2 + r + 1, cx2 + r + 1), off(cy2 - r, cx2 + r + 1), off(cy2 + r + 1, cx2 - r), off(cy2 - r, cx2 - r)};
            sTap[p] = pack_taps(o);
            const float side = (float)(2 * r + 1);
            sThr[p] = make_float2(row.thr, __fdiv_rn(1.0f, side * side));
        }
    }
    if (!FAST) {
        for (int p = threadIdx.x; p < a.P; p += GROUPS * TPG) {
            const PairRow row = load_pair(a.table, p);
            const int r = (int)row.r;
            const float side = (float)(2 * r + 1);
            sOff[p] = make_float4(row.oy1, row.ox1, row.oy2, row.ox2);
            sPar[p] = make_float4(row.thr, __fdiv_rn(1.0f, side * side), __int_as_float(r), 0.0f);
        }
    }
    __syncthreads();                             // barriers initialised, tables built

    const long long total = (long long)a.B * a.K;
    const long long stride = (long long)gridDim.x * GROUPS;
    unsigned int* wbuf = sWin + g * NBUF * G::GSTRIDE;
    const uint32_t bar0 = smem_u32(&bars[2 * g]);
    // thread 0 of the group launches the window load of keypoint k into buffer `buf`
    auto issue = [&](long long k, int buf) {
        float ky, kx;
        int z;
        if (t == 0 && win_needed(a, k, ky, kx, z)) {
            const int iy0 = (int)nearbyintf(fminf(ky, (float)(a.H - 1))), ix0 = (int)nearbyintf(fminf(fmaxf(kx, 0.0f), (float)(a.W - 1)));
            const uint32_t bar = bar0 + 8u * buf;
            mbar_arrive_expect_tx(bar, G::WR * G::WP * (int)sizeof(WT));
            tma_load_3d(smem_u32(wbuf + buf * G::GSTRIDE), &tmap, bar, (ix0 - HS + a.pad) & ~(G::EPW - 1), iy0 - HS + a.pad, z);
        }
    };
    long long kidx = (long long)blockIdx.x * GROUPS + g;
    if (kidx < total) issue(kidx, 0);
    uint32_t phase[2] = {0u, 0u};
    for (int n = 0; kidx < total; kidx += stride, ++n) {
        const int buf = NBUF == 2 ? (n & 1) : 0;
        if (NBUF == 2 && kidx + stride < total) issue(kidx + stride, buf ^ 1);   // that buffer was released by the barrier below
        float ky, kx;
        int z;
        if (win_needed(a, kidx, ky, kx, z)) {
            sparse_win_group<HS, ORIENTED, BILINEAR, NPP, WT>(a, kidx, z, ky, kx, reinterpret_cast<const WT*>(wbuf + buf * G::GSTRIDE),
                                                              bar0 + 8u * buf, phase[buf], red, sTheta, sTap, sThr, sIdx, sOff, sPar, g, t);
            phase[buf] ^= 1u;
        } else if (!(ky >= 0.0f)) {                                 // bad.py:461, :570 -> the row is all zeros
            float* out = a.desc + (size_t)kidx * a.P;
            for (int p = t; p < a.P; p += TPG) out[p] = 0.0f;
        }                                                           // else: non-integer image, sparse_bad_kernel does it
        group_bar(g);                                               // everyone is done with this buffer
        if (NBUF == 1 && kidx + stride < total) issue(kidx + stride, 0);
    }
}

template <int HS, int GROUPS, bool ORIENTED, bool BILINEAR, typename WT>
int launch_sparse_win(const SparseArgs& a, const void* I, cudaStream_t st) {
    using G = WinGeom<HS, WT>;
    const int Hi = a.H + 2 * a.pad + 1, IP = sizeof(WT) == 2 ? ipitch16(a.W, a.pad) : ipitch(a.W, a.pad);
    CUtensorMap tmap;
    OM_TRY(make_tmap_3d(&tmap, false, I, (uint64_t)IP, (uint64_t)Hi, (uint64_t)a.B, (uint64_t)IP, G::WP, G::WR, (int)sizeof(WT)));
    constexpr int NBUF = 1;
    const size_t smem = (size_t)GROUPS * NBUF * G::GSTRIDE * 4 +
                        ((!ORIENTED && !BILINEAR) ? (size_t)a.P * (sizeof(uint4) + sizeof(float2) + sizeof(unsigned short))
                                                  : (size_t)a.P * 2 * sizeof(float4));
    const long long total = (long long)a.B * a.K;
    const long long nblk = (total + GROUPS - 1) / GROUPS;
    const int per_sm = (int)(228 * 1024 / (smem + 1024));          // resident CTAs per SM by shared memory (228 KB, 1 KB reserved per CTA)
    const long long resident = 148ll * (per_sm < 1 ? 1 : per_sm);
    const unsigned grid = (unsigned)(nblk < resident ? nblk : resident);
    auto go = [&](auto kernel) -> int {
        OM_TRY(set_smem(kernel, smem));
        kernel<<<grid, GROUPS * TPG, smem, st>>>(tmap, a);
        return OM_OK;
    };
    if (a.P <= 4 * TPG) OM_TRY(go(sparse_win_kernel<HS, GROUPS, ORIENTED, BILINEAR, NBUF, 4, WT>));   // pairs per thread
    else OM_TRY(go(sparse_win_kernel<HS, GROUPS, ORIENTED, BILINEAR, NBUF, 8, WT>));
    OM_AFTER_LAUNCH();
    return OM_OK;
}

// ------------------------------------------------------------------------------------------
// orientation map for the whole image (module API of AngleEstimator)
// ------------------------------------------------------------------------------------------
constexpr int ATH = 16, ATW = 64;

__global__ void __launch_bounds__(256) angle_map_kernel(const float* image, int H, int W, const float* moments, int ps,
                                                        float* out) {
    extern __shared__ float sm[];
    const int half = ps / 2;
    const int IH = ATH + 2 * half, IW = ATW + 2 * half, IP = IW | 1;
    float* sW = sm;                 // 2*ps*ps
    float* sI = sm + 2 * ps * ps;   // IH x IP, zero padded (angle_estimation.py:161)
    const int z = blockIdx.z, ty0 = blockIdx.y * ATH, tx0 = blockIdx.x * ATW;
    const float* img = image + (size_t)z * H * W;
    for (int i = threadIdx.x; i < 2 * ps * ps; i += 256) sW[i] = moments[i];
    for (int i = threadIdx.x; i < IH * IW; i += 256) {
        const int ly = i / IW, lx = i % IW;
        const int gy = ty0 - half + ly, gx = tx0 - half + lx;
        sI[ly * IP + lx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? img[(size_t)gy * W + gx] : 0.0f;
    }
    __syncthreads();
    const int x = threadIdx.x % ATW, yq = threadIdx.x / ATW;   // 4 row groups, each thread 4 rows: yq, yq+4, ...
    float m10[4] = {0, 0, 0, 0}, m01[4] = {0, 0, 0, 0};
    for (int j = 0; j < ps; ++j) {
        for (int i = 0; i < ps; ++i) {
            const float wx = sW[j * ps + i], wy = sW[ps * ps + j * ps + i];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float v = sI[(yq + 4 * q + j) * IP + x + i];
                m10[q] = fmaf(wx, v, m10[q]);
                m01[q] = fmaf(wy, v, m01[q]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int gy = ty0 + yq + 4 * q, gx = tx0 + x;
        if (gy < H && gx < W) out[(size_t)z * H * W + (size_t)gy * W + gx] = atan2f(m01[q], m10[q]);   // :170
    }
}

// ------------------------------------------------------------------------------------------
// dense BAD: float32 integral image exactly as the reference builds it
// ------------------------------------------------------------------------------------------
// Integral images of the replicate-padded image, two flavours of the same two passes:
//   EXACT = false (dense path): the reference's float32 integral -- cumsum along H with a double
//           accumulator rounded to f32 per element, then cumsum along W of THAT (bad.py:70-72).
//   EXACT = true (sparse path): exact uint32 integral (wrap-around is harmless: box sums are
//           differences).  Exact only for integer-valued pixels in [0, 65535]; any other pixel raises
//           flags[image] and that image's keypoints are handled by the fp64 per-keypoint kernel.
// I is (H+2*pad+1) x ipitch with the zero row / column of bad.py:72; ipitch % 4 == 0 so that a
// keypoint's window is one TMA box.

// pass 1 (columns).  A column is cut into PC_SEG segments: every thread first sums its segment, the
// segment totals are combined in order through shared memory, then the thread re-walks its segment
// (an L1/L2 hit) from that base.  Double sums of <= 2^12 float32 values are exact for integer-valued
// images and within 1e-13 relative otherwise, i.e. identical after the single rounding.
constexpr int PC_COLS = 32, PC_SEG = 32;

template <bool EXACT, int LMAX>   // LMAX > 0: a segment (<= LMAX rows) is kept in registers between the two walks
__global__ void __launch_bounds__(PC_COLS * PC_SEG) prefix_cols_kernel(const float* image, int H, int W, int pad, void* Tout,
                                                                       unsigned int* flags, const unsigned int* gate) {
    if (gate != nullptr && *gate == 0u) return;              // fallback build: only when the banded build flagged an image
    using Acc = typename std::conditional<EXACT, unsigned int, double>::type;
    using Out = typename std::conditional<EXACT, unsigned int, float>::type;
    __shared__ Acc seg_total[PC_SEG][PC_COLS];
    const int Wp = W + 2 * pad, Hp = H + 2 * pad;
    const int cx = threadIdx.x % PC_COLS, sg = threadIdx.x / PC_COLS;
    const int xx = blockIdx.x * PC_COLS + cx;
    const int z = blockIdx.y;
    const int L = (Hp + PC_SEG - 1) / PC_SEG;
    const int y0 = sg * L, y1 = min(y0 + L, Hp);
    const bool live = xx < Wp;
    const float* col = image + (size_t)z * H * W + clampi(xx - pad, 0, W - 1);
    float vals[LMAX > 0 ? LMAX : 1];
    Acc acc = 0;
    bool odd = false;
    if (live) {
        if (LMAX > 0) {
#pragma unroll
            for (int k = 0; k < LMAX; ++k)
                vals[k] = (y0 + k < y1) ? __ldg(col + (size_t)clampi(y0 + k - pad, 0, H - 1) * W) : 0.0f;
#pragma unroll
            for (int k = 0; k < LMAX; ++k) {
                if (EXACT) {
                    odd |= !(vals[k] == rintf(vals[k]) && vals[k] >= 0.0f && vals[k] <= 65535.0f);
                    acc += (Acc)__float2uint_rz(vals[k]);
                } else {
                    acc += (Acc)vals[k];
                }
            }
        } else {
#pragma unroll 4
            for (int yy = y0; yy < y1; ++yy) {
                const float v = __ldg(col + (size_t)clampi(yy - pad, 0, H - 1) * W);
                if (EXACT) {
                    odd |= !(v == rintf(v) && v >= 0.0f && v <= 65535.0f);
                    acc += (Acc)__float2uint_rz(v);
                } else {
                    acc += (Acc)v;
                }
            }
        }
    }
    seg_total[sg][cx] = acc;
    const int any_odd = __syncthreads_or(odd ? 1 : 0);
    if (EXACT && any_odd && threadIdx.x == 0) {
        atomicOr(&flags[z], 1u);
        atomicOr(&flags[gridDim.y], 1u);                     // "some image of the batch is flagged"
    }
    if (!live) return;
    acc = 0;
    for (int q = 0; q < sg; ++q) acc += seg_total[q][cx];
    Out* dst = reinterpret_cast<Out*>(Tout) + (size_t)z * Hp * Wp + xx;
    if (LMAX > 0) {
#pragma unroll
        for (int k = 0; k < LMAX; ++k) {
            if (y0 + k < y1) {
                acc += EXACT ? (Acc)__float2uint_rz(vals[k]) : (Acc)vals[k];
                dst[(size_t)(y0 + k) * Wp] = (Out)acc;
            }
        }
    } else {
#pragma unroll 4
        for (int yy = y0; yy < y1; ++yy) {
            const float v = __ldg(col + (size_t)clampi(yy - pad, 0, H - 1) * W);
            acc += EXACT ? (Acc)__float2uint_rz(v) : (Acc)v;
            dst[(size_t)yy * Wp] = (Out)acc;
        }
    }
}

// pass 2 (rows).  One warp per row: the row is staged in shared memory, lane l sums the l-th segment,
// the 32 segment totals are scanned with shuffles, then every lane re-walks its segment and the row is
// written back coalesced, shifted by the zero column.
constexpr int IR_ROWS = 8;

template <bool EXACT>
__global__ void __launch_bounds__(IR_ROWS * 32) prefix_rows_kernel(const void* Tin, int H, int W, int pad, int SL, void* Iout,
                                                                   const unsigned int* gate) {
    if (gate != nullptr && *gate == 0u) return;
    using Acc = typename std::conditional<EXACT, unsigned int, double>::type;
    using Val = typename std::conditional<EXACT, unsigned int, float>::type;
    extern __shared__ __align__(16) unsigned char sRowRaw[];
    Val* sRow = reinterpret_cast<Val*>(sRowRaw);             // IR_ROWS x (32*SL)
    const int Wp = W + 2 * pad, Hp = H + 2 * pad, IP = ipitch(W, pad);
    const int z = blockIdx.y;
    const int wrp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * IR_ROWS + wrp;              // row of I, 0 .. Hp
    if (row > Hp) return;
    Val* dst = reinterpret_cast<Val*>(Iout) + (size_t)z * (Hp + 1) * IP + (size_t)row * IP;
    if (row == 0) {
        for (int x = lane; x < IP; x += 32) dst[x] = 0;
        return;
    }
    const Val* src = reinterpret_cast<const Val*>(Tin) + (size_t)z * Hp * Wp + (size_t)(row - 1) * Wp;
    Val* buf = sRow + (size_t)wrp * 32 * SL;
    for (int x = lane; x < 32 * SL; x += 32) buf[x] = x < Wp ? __ldg(src + x) : (Val)0;
    __syncwarp();
    Val* seg = buf + lane * SL;                              // SL is odd: conflict-free
    Acc tot = 0;
    for (int k = 0; k < SL; ++k) tot += (Acc)seg[k];
    Acc incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const Acc nb = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += nb;
    }
    Acc acc = incl - tot;                                    // exclusive base of this segment
    for (int k = 0; k < SL; ++k) {
        acc += (Acc)seg[k];
        seg[k] = (Val)acc;
    }
    __syncwarp();
    if (lane == 0) dst[0] = 0;
    for (int x = lane; x < Wp; x += 32) dst[x + 1] = buf[x];
    for (int x = Wp + 1 + lane; x < IP; x += 32) dst[x] = 0;
}

template <bool EXACT>
int build_prefix(const float* image, int B, int H, int W, int pad, void* T, void* I, unsigned int* flags, cudaStream_t st,
                 const unsigned int* gate = nullptr) {
    const int Wp = W + 2 * pad, Hp = H + 2 * pad;
    const dim3 cgrid((Wp + PC_COLS - 1) / PC_COLS, B);
    const int L = (Hp + PC_SEG - 1) / PC_SEG;
    if (EXACT && L <= 20) prefix_cols_kernel<EXACT, 20><<<cgrid, PC_COLS * PC_SEG, 0, st>>>(image, H, W, pad, T, flags, gate);
    else prefix_cols_kernel<EXACT, 0><<<cgrid, PC_COLS * PC_SEG, 0, st>>>(image, H, W, pad, T, flags, gate);
    OM_AFTER_LAUNCH();
    const int SL = ((Wp + 31) / 32) | 1;
    const size_t smem = (size_t)IR_ROWS * 32 * SL * 4;
    if (smem > 200 * 1024) return OM_ERR_LIMIT;              // W <= ~6300
    OM_TRY(set_smem(prefix_rows_kernel<EXACT>, smem));
    prefix_rows_kernel<EXACT><<<dim3((Hp + 1 + IR_ROWS - 1) / IR_ROWS, B), IR_ROWS * 32, smem, st>>>(T, H, W, pad, SL, I, gate);
    OM_AFTER_LAUNCH();
    return OM_OK;
}

// ------------------------------------------------------------------------------------------
// banded integral image (the default build; the two prefix passes above remain for float-valued images on the dense
// path and for very wide images)
// ------------------------------------------------------------------------------------------
// The image is read in its own type (float32 or uint8) and the integral is exact uint32 arithmetic, stored as uint32
// (sparse path: box sums are wrap-around differences) or as float32 = RN(exact) (dense path: exactly what the
// reference's double-accumulated float32 cumsums produce for integer-valued pixels, bad.py:70-72).
//   pass A  band_colsum_kernel: the padded image is cut into bands of IB_ROWS rows; one CTA per (band, image) sums every
//           column of its band (tiny output: bands x pitch words per image) and raises the per-image flag when a pixel
//           is not an integer in [0, maxv] (float input only).
//   pass B  integral_band_kernel: one CTA per (band, image) owns ALL columns of its band: thread t holds output columns
//           8t..8t+7.  It starts from the integral row above the band (column sums of the bands above, scanned across
//           the CTA once), then walks its rows in groups of IB_G: row prefix = 8-element local prefix + warp scan
//           (shuffles) + warp totals through shared memory (ONE block barrier per group, double buffered), vertical
//           accumulation in registers (one 3-input add per pixel), two 16-byte stores per row.  Two groups of rows are in
//           flight while a third is scanned.  The image is read twice (pass A, pass B) and the integral written once: no
//           intermediate integral-sized array exists (the two-pass build wrote and re-read one), and every (band, image)
//           is independent (no look-back chain).
// float -> integer conversions use the 2^23 trick (v + 2^23 holds the integer v in its low mantissa bits for
// 0 <= v < 2^23; (v + 2^23) - 2^23 == v exactly when v is such an integer): no conversion instructions, no branches.
int g_band_rows = 32;              // padded-image rows per band (om_debug_band_rows)
int g_sparse_i16 = 1;              // sparse path: integral stored modulo 2^16 (om_debug_band_rows(-16) / (-32) switch it on / off)
constexpr int IB_G = 2;            // rows per group
constexpr int IB_PX = 8;           // output columns per thread
constexpr float IB_MAGIC = 8388608.0f;

// which pixels thread t's output columns c = 8t + j read: padded pixel column c - 1 (column 0 of the integral and the
// columns beyond the padded width hold zero), replicate-clamped to the image
// ALIGNED images (W % 8 == 0, 8-pixel aligned base; (1 + pad) % 8 == 0 for every pad in use): a thread's eight pixels are
// either eight consecutive aligned pixels of the image (vector loads) or all the same replicated border pixel (one load)
// -- no thread straddles a border.  Other images take clamped scalar loads (separate instantiation).
struct BandCols {
    unsigned int in;    // bit j: output column 8t + j is a pixel column
    bool vec;           // the eight pixels are consecutive, unclamped and aligned: vector loads
    int x0;             // image column of j = 0 (before clamping)
    int wmax;           // W - 1
};
__device__ __forceinline__ BandCols band_cols(int t, int W, int pad, bool aligned_rows) {
    BandCols b;
    const int Wp = W + 2 * pad;
    b.in = 0u;
#pragma unroll
    for (int j = 0; j < IB_PX; ++j) {
        const int c = IB_PX * t + j;
        if (c >= 1 && c <= Wp) b.in |= 1u << j;
    }
    b.x0 = IB_PX * t - 1 - pad;
    b.wmax = W - 1;
    b.vec = aligned_rows && b.x0 >= 0 && b.x0 + IB_PX - 1 <= W - 1 && (b.x0 & (IB_PX - 1)) == 0;
    return b;
}
// raw pixels of one row (loads only: nothing here waits for the data)
template <typename TIn, bool ALIGNED>
__device__ __forceinline__ void band_fetch(const TIn* rowp, const BandCols& b, TIn (&raw)[IB_PX]) {
    if (ALIGNED && !b.vec) {                                    // wholly inside the left or the right padding
        const TIn v = __ldg(rowp + clampi(b.x0, 0, b.wmax));
#pragma unroll
        for (int j = 0; j < IB_PX; ++j) raw[j] = v;
        return;
    }
    if (ALIGNED) {
        if constexpr (sizeof(TIn) == 4) {
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(rowp + b.x0)), q1 = __ldg(reinterpret_cast<const float4*>(rowp + b.x0 + 4));
            raw[0] = q0.x; raw[1] = q0.y; raw[2] = q0.z; raw[3] = q0.w; raw[4] = q1.x; raw[5] = q1.y; raw[6] = q1.z; raw[7] = q1.w;
        } else {
            const uint2 q = __ldg(reinterpret_cast<const uint2*>(rowp + b.x0));
            raw[0] = (TIn)(q.x & 0xFFu); raw[1] = (TIn)((q.x >> 8) & 0xFFu); raw[2] = (TIn)((q.x >> 16) & 0xFFu); raw[3] = (TIn)(q.x >> 24);
            raw[4] = (TIn)(q.y & 0xFFu); raw[5] = (TIn)((q.y >> 8) & 0xFFu); raw[6] = (TIn)((q.y >> 16) & 0xFFu); raw[7] = (TIn)(q.y >> 24);
        }
    } else {
#pragma unroll
        for (int j = 0; j < IB_PX; ++j) raw[j] = __ldg(rowp + clampi(b.x0 + j, 0, b.wmax));   // clamped: always a valid address
    }
}
template <typename TIn>
__device__ __forceinline__ unsigned int band_uint(TIn v) {
    if constexpr (sizeof(TIn) == 4) return __float_as_uint(__fadd_rn(v, IB_MAGIC)) & 0x007FFFFFu;
    else return (unsigned int)v;
}

template <typename TIn, int NT, bool ALIGNED>
__global__ void __launch_bounds__(NT, (NT <= 128 ? 8 : (NT <= 256 ? 4 : 1)))
band_colsum_kernel(const TIn* image, int H, int W, int pad, int nb, int IB_ROWS, int CP, float maxv, unsigned int* colsum,
                   unsigned int* flags) {
    const int z = blockIdx.y, k = blockIdx.x, t = threadIdx.x;
    const int Hp = H + 2 * pad;
    const bool live = IB_PX * t < CP;
    const BandCols bc = band_cols(t, W, pad, ALIGNED);
    const TIn* img = image + (size_t)z * H * W;
    const int y0 = k * IB_ROWS, y1 = min(y0 + IB_ROWS, Hp);
    unsigned int s[IB_PX];
#pragma unroll
    for (int j = 0; j < IB_PX; ++j) s[j] = 0u;
    bool odd = false;
    float mn = 0.0f, mx = 0.0f;
    if (live) {
        constexpr int RA = 4;                                   // rows in flight per thread
        for (int y = y0; y < y1; y += RA) {
            TIn raw[RA][IB_PX];
#pragma unroll
            for (int g = 0; g < RA; ++g) band_fetch<TIn, ALIGNED>(img + (size_t)clampi(min(y + g, y1 - 1) - pad, 0, H - 1) * W, bc, raw[g]);
#pragma unroll
            for (int g = 0; g < RA; ++g) {
                if (y + g < y1) {                               // block-uniform
#pragma unroll
                    for (int j = 0; j < IB_PX; ++j) {
                        if constexpr (sizeof(TIn) == 4) {
                            const float v = raw[g][j];
                            const float tt = __fadd_rn(v, IB_MAGIC);
                            odd |= __fsub_rn(tt, IB_MAGIC) != v;                  // not an integer (or NaN, or beyond 2^23)
                            mn = fminf(mn, v);
                            mx = fmaxf(mx, v);
                            s[j] += __float_as_uint(tt) & 0x007FFFFFu;
                        } else {
                            s[j] += (unsigned int)raw[g][j];
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < IB_PX; ++j)
            if (!((bc.in >> j) & 1u)) s[j] = 0u;
        unsigned int* dst = colsum + ((size_t)z * nb + k) * CP + IB_PX * t;
        *reinterpret_cast<uint4*>(dst) = make_uint4(s[0], s[1], s[2], s[3]);
        *reinterpret_cast<uint4*>(dst + 4) = make_uint4(s[4], s[5], s[6], s[7]);
    }
    if (sizeof(TIn) == 4) {
        odd = odd || mn < 0.0f || mx > maxv;
        const int any_odd = __syncthreads_or(odd ? 1 : 0);
        if (any_odd && t == 0) {
            atomicOr(&flags[z], 1u);
            atomicOr(&flags[gridDim.y], 1u);                 // "some image of the batch is flagged"
        }
    }
}

template <typename TIn, typename TOut, int NT, bool ALIGNED>
__global__ void __launch_bounds__(NT, (NT <= 128 ? 7 : (NT <= 256 ? 3 : 1)))
integral_band_kernel(const TIn* image, int H, int W, int pad, int nb, int IB_ROWS, int IP, int CP, const unsigned int* colsum,
                     TOut* Iout) {
    constexpr int NWARP = NT / 32;
    __shared__ __align__(16) unsigned int wtot[2][NWARP][4];    // [buffer][warp][row of the group (padded to 4)]
    const int z = blockIdx.y, k = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    const bool live = IB_PX * t < IP;
    const bool live_hi = IB_PX * t + 4 < IP;                    // ipitch is a multiple of 4, not of 8
    const BandCols bc = band_cols(t, W, pad, ALIGNED);
    // beyond the padded width the integral holds zero (never a tap); column 0 is zero by construction (prefix of nothing)
    const unsigned int first_zero = (unsigned int)max(0, min(IB_PX, Wp + 1 - IB_PX * t));   // columns j >= first_zero are stored as 0
    const TIn* img = image + (size_t)z * H * W;
    TOut* Iz = Iout + (size_t)z * (Hp + 1) * IP;
    const int y0 = k * IB_ROWS, y1 = min(y0 + IB_ROWS, Hp);
    const unsigned full = 0xffffffffu;
    auto put = [&](int irow, const unsigned int (&a)[IB_PX]) {  // integral row irow, this thread's eight columns
        if (!live) return;
        unsigned int o[IB_PX];
#pragma unroll
        for (int j = 0; j < IB_PX; ++j) o[j] = a[j];
        if (first_zero < (unsigned int)IB_PX) {                  // the thread(s) at the right end of the padded width only
#pragma unroll
            for (int j = 0; j < IB_PX; ++j) o[j] = (unsigned int)j < first_zero ? a[j] : 0u;
        }
        TOut* dst = Iz + (size_t)irow * IP + IB_PX * t;
        if constexpr (sizeof(TOut) == 2) {                      // the integral modulo 2^16 (IP is a multiple of eight here)
            *reinterpret_cast<uint4*>(dst) = make_uint4((o[0] & 0xFFFFu) | (o[1] << 16), (o[2] & 0xFFFFu) | (o[3] << 16),
                                                        (o[4] & 0xFFFFu) | (o[5] << 16), (o[6] & 0xFFFFu) | (o[7] << 16));
        } else if constexpr (std::is_same<TOut, float>::value) {
            *reinterpret_cast<float4*>(dst) = make_float4(__uint2float_rn(o[0]), __uint2float_rn(o[1]), __uint2float_rn(o[2]), __uint2float_rn(o[3]));
            if (live_hi)
                *reinterpret_cast<float4*>(dst + 4) = make_float4(__uint2float_rn(o[4]), __uint2float_rn(o[5]), __uint2float_rn(o[6]), __uint2float_rn(o[7]));
        } else {
            *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
            if (live_hi) *reinterpret_cast<uint4*>(dst + 4) = make_uint4(o[4], o[5], o[6], o[7]);
        }
    };
    // exclusive prefix over the CTA of one value per thread, for the IB_G rows of a group at once
    int buf = 0;
    auto cta_excl = [&](const unsigned int (&tot)[IB_G], unsigned int (&excl)[IB_G]) {
        unsigned int incl[IB_G];
#pragma unroll
        for (int g = 0; g < IB_G; ++g) incl[g] = tot[g];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int g = 0; g < IB_G; ++g) {
                const unsigned int n = __shfl_up_sync(full, incl[g], o);
                incl[g] += lane >= o ? n : 0u;
            }
        }
        if (lane == 31) {
#pragma unroll
            for (int g = 0; g < IB_G; ++g) wtot[buf][warp][g] = incl[g];
        }
        __syncthreads();
        unsigned int wb[IB_G];
#pragma unroll
        for (int g = 0; g < IB_G; ++g) wb[g] = 0u;
#pragma unroll
        for (int w = 0; w < NWARP - 1; ++w) {                   // totals of the warps before this one
            const uint2 v = *reinterpret_cast<const uint2*>(&wtot[buf][w][0]);
            if (w < warp) { wb[0] += v.x; if (IB_G > 1) wb[IB_G - 1] += v.y; }
        }
#pragma unroll
        for (int g = 0; g < IB_G; ++g) excl[g] = wb[g] + incl[g] - tot[g];
        buf ^= 1;
    };
    static_assert(IB_G == 2, "the warp-total exchange above moves two rows per load");

    // integral row above the band: row prefix of the column sums of the bands above
    unsigned int acc[IB_PX];
    {
        unsigned int base[IB_PX];
#pragma unroll
        for (int j = 0; j < IB_PX; ++j) base[j] = 0u;
        if (live) {
            const unsigned int* cs = colsum + (size_t)z * nb * CP + IB_PX * t;
#pragma unroll 4
            for (int q = 0; q < k; ++q) {
                const uint4 c0 = __ldg(reinterpret_cast<const uint4*>(cs + (size_t)q * CP)), c1 = __ldg(reinterpret_cast<const uint4*>(cs + (size_t)q * CP + 4));
                base[0] += c0.x; base[1] += c0.y; base[2] += c0.z; base[3] += c0.w;
                base[4] += c1.x; base[5] += c1.y; base[6] += c1.z; base[7] += c1.w;
            }
        }
#pragma unroll
        for (int j = 1; j < IB_PX; ++j) base[j] += base[j - 1];
        unsigned int tot[IB_G] = {base[IB_PX - 1], 0u}, excl[IB_G];
        cta_excl(tot, excl);
#pragma unroll
        for (int j = 0; j < IB_PX; ++j) acc[j] = excl[0] + base[j];
    }
    if (k == 0) {
        unsigned int zero[IB_PX];
#pragma unroll
        for (int j = 0; j < IB_PX; ++j) zero[j] = 0u;
        put(0, zero);                                            // the zero row of bad.py:72
    }
    auto fetch = [&](int y, TIn (&dst)[IB_G][IB_PX]) {
#pragma unroll
        for (int g = 0; g < IB_G; ++g)      // (threads beyond the pitch have in == 0: their clamped loads are masked to zero)
            band_fetch<TIn, ALIGNED>(img + (size_t)clampi(min(y + g, y1 - 1) - pad, 0, H - 1) * W, bc, dst[g]);
    };
    auto scan_rows = [&](int y, const TIn (&raw)[IB_G][IB_PX]) {
        unsigned int v[IB_G][IB_PX], tot[IB_G], excl[IB_G];
#pragma unroll
        for (int g = 0; g < IB_G; ++g) {
#pragma unroll
            for (int j = 0; j < IB_PX; ++j) v[g][j] = band_uint<TIn>(raw[g][j]);
            if (bc.in != 0xFFu) {                               // edge threads only: column 0, columns beyond the padded width
#pragma unroll
                for (int j = 0; j < IB_PX; ++j) v[g][j] = ((bc.in >> j) & 1u) ? v[g][j] : 0u;
            }
#pragma unroll
            for (int j = 1; j < IB_PX; ++j) v[g][j] += v[g][j - 1];      // inclusive prefix of the thread's eight pixels
            tot[g] = v[g][IB_PX - 1];
        }
        cta_excl(tot, excl);
#pragma unroll
        for (int g = 0; g < IB_G; ++g) {
            if (y + g < y1) {                                   // block-uniform
#pragma unroll
                for (int j = 0; j < IB_PX; ++j) acc[j] = acc[j] + v[g][j] + excl[g];
                put(y + g + 1, acc);
            }
        }
    };
    // three ring slots, rotated by name so that every index stays a compile-time constant
    TIn ga[IB_G][IB_PX], gb[IB_G][IB_PX], gc[IB_G][IB_PX];
    fetch(y0, ga);
    fetch(y0 + IB_G, gb);
    for (int y = y0; y < y1; y += 3 * IB_G) {
        fetch(y + 2 * IB_G, gc);
        scan_rows(y, ga);
        if (y + IB_G >= y1) break;
        fetch(y + 3 * IB_G, ga);
        scan_rows(y + IB_G, gb);
        if (y + 2 * IB_G >= y1) break;
        fetch(y + 4 * IB_G, gb);
        scan_rows(y + 2 * IB_G, gc);
    }
}

inline int band_colsum_pitch(int W, int pad) { return (ipitch(W, pad) + IB_PX - 1) / IB_PX * IB_PX; }

// ipitch <= 8 * NT.  Returns OM_ERR_LIMIT when the image is too wide for one CTA per band (the caller falls back).
template <typename TIn, typename TOut>
int build_integral_banded(const TIn* image, int B, int H, int W, int pad, float maxv, unsigned int* colsum, TOut* I,
                          unsigned int* flags, cudaStream_t st) {
    const int Hp = H + 2 * pad, IP = sizeof(TOut) == 2 ? ipitch16(W, pad) : ipitch(W, pad), CP = band_colsum_pitch(W, pad);
    const int IB_ROWS = g_band_rows;
    const int nb = (Hp + IB_ROWS - 1) / IB_ROWS;
    if (CP > IB_PX * 1024 || B > 65535) return OM_ERR_LIMIT;
    const dim3 grid(nb, B);
    const bool aligned = (W & 7) == 0 && ((1 + pad) & 7) == 0 && (reinterpret_cast<uintptr_t>(image) & (IB_PX * sizeof(TIn) - 1)) == 0;
#define OM_BAND_LAUNCH(NT)                                                                                              \
    do {                                                                                                                \
        if (aligned) band_colsum_kernel<TIn, NT, true><<<grid, NT, 0, st>>>(image, H, W, pad, nb, IB_ROWS, CP, maxv, colsum, flags);      \
        else band_colsum_kernel<TIn, NT, false><<<grid, NT, 0, st>>>(image, H, W, pad, nb, IB_ROWS, CP, maxv, colsum, flags);             \
        OM_AFTER_LAUNCH();                                                                                              \
        if (aligned) integral_band_kernel<TIn, TOut, NT, true><<<grid, NT, 0, st>>>(image, H, W, pad, nb, IB_ROWS, IP, CP, colsum, I);    \
        else integral_band_kernel<TIn, TOut, NT, false><<<grid, NT, 0, st>>>(image, H, W, pad, nb, IB_ROWS, IP, CP, colsum, I);           \
        OM_AFTER_LAUNCH();                                                                                              \
    } while (0)
    if (CP <= IB_PX * 64) OM_BAND_LAUNCH(64);
    else if (CP <= IB_PX * 96) OM_BAND_LAUNCH(96);
    else if (CP <= IB_PX * 128) OM_BAND_LAUNCH(128);
    else if (CP <= IB_PX * 256) OM_BAND_LAUNCH(256);
    else if (CP <= IB_PX * 512) OM_BAND_LAUNCH(512);
    else OM_BAND_LAUNCH(1024);
#undef OM_BAND_LAUNCH
    return OM_OK;
}
inline size_t band_colsum_bytes(int B, int H, int W, int pad) {
    const int nb = (H + 2 * pad + 8 - 1) / 8;                   // sized for the smallest band height (8 rows)
    return align_up((size_t)B * nb * band_colsum_pitch(W, pad) * sizeof(unsigned int));
}

// box mean of the reference's dense path at padded-integral centre (cy,cx) (already +MAXR):
// bad.py:87-99, fp32 taps in the reference's order, fp32 divide by the area
__device__ __forceinline__ float dense_box_mean(const float* I, int pitch, int cy, int cx, int r) {
    const int y0 = cy - r, x0 = cx - r, y1 = cy + r + 1, x1 = cx + r + 1;
    const float s = __fadd_rn(__fsub_rn(__fsub_rn(I[y1 * pitch + x1], I[y0 * pitch + x1]), I[y1 * pitch + x0]),
                              I[y0 * pitch + x0]);
    const float side = (float)(2 * r + 1);
    return __fdiv_rn(s, side * side);                        // bad.py:99
}

// planes[r][y][x] = box mean of radius r centred on image pixel (y,x), r = 0..7
__global__ void __launch_bounds__(128) box_planes_kernel(const float* I, int H, int W, float* planes) {
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int y = blockIdx.y;
    const int z = blockIdx.z;
    if (x >= W) return;
    const int pitch = ipitch(W, MAXR);
    const float* Iz = I + (size_t)z * (H + 2 * MAXR + 1) * pitch;
    float* pz = planes + (size_t)z * (MAXR + 1) * H * W + (size_t)y * W + x;
#pragma unroll
    for (int r = 0; r <= MAXR; ++r) pz[(size_t)r * H * W] = dense_box_mean(Iz, pitch, y + MAXR, x + MAXR, r);
}

__global__ void __launch_bounds__(128) dense_map_kernel(const float* planes, int H, int W, const float* table, int P,
                                                        int mode, float temperature, float* out) {
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int y = blockIdx.y;
    const int z = blockIdx.z;
    if (x >= W) return;
    const float* pz = planes + (size_t)z * (MAXR + 1) * H * W;
    float* oz = out + (size_t)z * P * H * W + (size_t)y * W + x;
    for (int p = 0; p < P; ++p) {
        const PairRow row = load_pair(table, p);
        const int r = (int)row.r;
        const float* pl = pz + (size_t)r * H * W;
        // bad.py:81-82: centre = clamp(base + offset, 0, dim-1)
        const int y1 = clampi(y + (int)row.oy1, 0, H - 1), x1 = clampi(x + (int)row.ox1, 0, W - 1);
        const int y2 = clampi(y + (int)row.oy2, 0, H - 1), x2 = clampi(x + (int)row.ox2, 0, W - 1);
        const float diff = __fsub_rn(__ldg(pl + (size_t)y1 * W + x1), __ldg(pl + (size_t)y2 * W + x2));
        oz[(size_t)p * H * W] = finish_value(diff, row.thr, mode, temperature);
    }
}

struct DenseKpArgs {
    const float* I;            // (B, H+15, W+15) float32 integral
    int B, H, W;
    const float* kpts;
    int K;
    const float* table;
    int P, mode;
    float temperature;
    int normalize;
    float* desc;
    int dbg;                   // diagnosis: bit 1 = skip the window fetch, bit 2 = skip the pair arithmetic
};

// Dense-path descriptors at the K keypoints only: bilinear blend of the dense map's values at the
// four neighbouring pixels (bad.py:277-333), mask, L2 normalise (shi_tomasi_bad_sinkhorn.py:143-158, :213-214).
// One 64-thread group per keypoint stages the 47x47 window of the float32 integral that covers every
// tap of the four neighbours.  The bilinear weights are the same for all pairs of a keypoint and are
// exactly 0 for the neighbours an exact coordinate round trip does not touch (0 * value adds
// nothing), so those neighbours are skipped.  Interior keypoints (no centre clamps, bad.py:81-82)
// read their 8 taps per pair at window offsets precomputed once per CTA.
// 47 rows x 47 columns are needed; the TMA box must start on a 16-byte boundary in global memory, so it starts up
// to 3 columns early: 48 rows x 52 columns
constexpr int DK_LO = 15, DK_ROWS = 48, DK_SPAN = 52;

// bilinear sampling position of a keypoint (bad.py:311-312: kp / (dim-1+1e-8) * 2 - 1, then ATen unnormalise + clip)
struct DenseKp {
    int iy0, ix0;          // top-left neighbour
    float w, s;            // fractional parts along x and y
};
__device__ __forceinline__ DenseKp dense_kp_pos(float ky, float kx, int H, int W) {
    const float yc = fminf(fmaxf(ky, 0.0f), (float)(H - 1));
    const float xc = fminf(fmaxf(kx, 0.0f), (float)(W - 1));
    const float gyn = __fsub_rn(__fmul_rn(__fdiv_rn(yc, (float)(H - 1 + 1e-8)), 2.0f), 1.0f);
    const float gxn = __fsub_rn(__fmul_rn(__fdiv_rn(xc, (float)(W - 1 + 1e-8)), 2.0f), 1.0f);
    const float uy = fminf(fmaxf(__fmul_rn(__fadd_rn(gyn, 1.0f), (float)(H - 1) * 0.5f), 0.0f), (float)(H - 1));
    const float ux = fminf(fmaxf(__fmul_rn(__fadd_rn(gxn, 1.0f), (float)(W - 1) * 0.5f), 0.0f), (float)(W - 1));
    const float fy = floorf(uy), fx = floorf(ux);
    DenseKp k;
    k.iy0 = (int)fy; k.ix0 = (int)fx; k.w = ux - fx; k.s = uy - fy;
    return k;
}

// one keypoint whose window `Lbox` is in flight / has landed on mbarrier `bar` (phase `parity`)
// value of one pair at a keypoint near the image border (centre clamps, bad.py:81-82): rare, kept out of line so that the
// two interior paths stay small
__device__ __noinline__ float dense_pair_border(const float* table, int pair, const float* L, int H, int W, int oy, int ox,
                                                int iy0, int ix0, float w0, float w1, float w2, float w3, int mode,
                                                float temperature) {
    constexpr int SPAN = DK_SPAN;
    const PairRow row = load_pair(table, pair);
    const int r = (int)row.r;
    const int ys = min(iy0 + 1, H - 1), xe = min(ix0 + 1, W - 1);
    const int py[4] = {iy0, iy0, ys, ys}, px[4] = {ix0, xe, ix0, xe};
    const float wgt[4] = {w0, w1, w2, w3};
    float v = 0.0f;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        if (wgt[k] != 0.0f) {
            const int c1y = clampi(py[k] + (int)row.oy1, 0, H - 1) + MAXR - oy, c1x = clampi(px[k] + (int)row.ox1, 0, W - 1) + MAXR - ox;
            const int c2y = clampi(py[k] + (int)row.oy2, 0, H - 1) + MAXR - oy, c2x = clampi(px[k] + (int)row.ox2, 0, W - 1) + MAXR - ox;
            const float diff = __fsub_rn(dense_box_mean(L, SPAN, c1y, c1x, r), dense_box_mean(L, SPAN, c2y, c2x, r));
            v += finish_value(diff, row.thr, mode, temperature) * wgt[k];
        }
    }
    return v;
}

// one keypoint whose window `Lbox` is in flight / has landed on mbarrier `bar` (phase `parity`).  NPP = pairs per thread
// (P / 64): the pair loops are fully unrolled over exactly NPP slots so that the NPP independent tap chains overlap.
template <bool CPA, int NPP>
__device__ __forceinline__ void dense_kp_group(const DenseKpArgs& a, long long kidx, float ky, float kx, const float* Lbox,
                                               uint32_t bar, uint32_t parity, float* red, const uint4* sTap,
                                               const float4* sThr, int g, int t) {
    constexpr int LO = DK_LO, SPAN = DK_SPAN;
    const int H = a.H, W = a.W;
    float* out = a.desc + (size_t)kidx * a.P;
    const DenseKp kp = dense_kp_pos(ky, kx, H, W);
    const int iy0 = kp.iy0, ix0 = kp.ix0;
    const float w = kp.w, e = 1.0f - w, s = kp.s, n = 1.0f - s;
    const int oy = iy0 - LO, ox = ix0 - LO;     // integral coords of L(0,0)
    const float wgt[4] = {n * e, n * w, s * e, s * w};                  // nw, ne, sw, se
    const float* L = Lbox + (ox & 3);           // window column 0 inside the 16-byte aligned box
    if (CPA) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");            // this thread's 16-byte copies have landed
        group_bar(g);                                                   // ... and everyone else's
    } else {
        mbar_wait(bar, parity);
    }

    float d[NPP];
    int pidx[NPP];
    float ss = 0.0f;
    // no centre clamp for any of the four neighbours <=> their offsets in [-15,14] stay inside the image
    const bool interior = iy0 >= 15 && iy0 + 1 + 14 <= H - 1 && ix0 >= 15 && ix0 + 1 + 14 <= W - 1;
    const char* lbytes = reinterpret_cast<const char*>(L);
    auto ldl = [&](unsigned int byte_off) -> float { return *reinterpret_cast<const float*>(lbytes + byte_off); };
    if (a.dbg & 4) {
#pragma unroll
        for (int q = 0; q < NPP; ++q) { d[q] = L[t + q * TPG]; pidx[q] = t + q * TPG < a.P ? t + q * TPG : -1; }
    } else if (interior && w == 0.0f && s == 0.0f) {
        // keypoint whose sampling position is exactly on the pixel grid: the weights are {1, 0, 0, 0}, one neighbour, no
        // per-neighbour tests; 0 + value * 1 of the general path is value (up to the sign of a zero)
#pragma unroll
        for (int q = 0; q < NPP; ++q) {
            const int p = t + q * TPG;                                  // slot; sThr[p].w holds its pair index
            d[q] = 0.0f; pidx[q] = -1;
            if (p >= a.P) continue;                                     // only when P is not a multiple of 64
            const uint4 tp = sTap[p];
            const float4 tb = sThr[p];
            const float s1 = __fadd_rn(__fsub_rn(__fsub_rn(ldl(tap_lo(tp.x)), ldl(tap_hi(tp.x))), ldl(tap_lo(tp.y))), ldl(tap_hi(tp.y)));
            const float s2 = __fadd_rn(__fsub_rn(__fsub_rn(ldl(tap_lo(tp.z)), ldl(tap_hi(tp.z))), ldl(tap_lo(tp.w))), ldl(tap_hi(tp.w)));
            const float diff = __fsub_rn(div_area(s1, tb.y, tb.z), div_area(s2, tb.y, tb.z));               // bad.py:99, :110
            d[q] = __fadd_rn(0.0f, finish_value(diff, tb.x, a.mode, a.temperature));
            pidx[q] = __float_as_int(tb.w);
        }
    } else if (interior) {
#pragma unroll
        for (int q = 0; q < NPP; ++q) {
            const int p = t + q * TPG;
            d[q] = 0.0f; pidx[q] = -1;
            if (p >= a.P) continue;
            const uint4 tp = sTap[p];
            const float4 tb = sThr[p];
            const unsigned a0 = tap_lo(tp.x), a1 = tap_hi(tp.x), a2 = tap_lo(tp.y), a3 = tap_hi(tp.y);
            const unsigned c0 = tap_lo(tp.z), c1 = tap_hi(tp.z), c2 = tap_lo(tp.w), c3 = tap_hi(tp.w);
            float v = 0.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (wgt[k] != 0.0f) {                                   // group-uniform
                    const unsigned int lk = ((k >> 1) * SPAN + (k & 1)) * 4;
                    const float s1 = __fadd_rn(__fsub_rn(__fsub_rn(ldl(lk + a0), ldl(lk + a1)), ldl(lk + a2)), ldl(lk + a3));
                    const float s2 = __fadd_rn(__fsub_rn(__fsub_rn(ldl(lk + c0), ldl(lk + c1)), ldl(lk + c2)), ldl(lk + c3));
                    const float diff = __fsub_rn(div_area(s1, tb.y, tb.z), div_area(s2, tb.y, tb.z));       // bad.py:99, :110
                    v += finish_value(diff, tb.x, a.mode, a.temperature) * wgt[k];
                }
            }
            d[q] = v;
            pidx[q] = __float_as_int(tb.w);
        }
    } else {
#pragma unroll
        for (int q = 0; q < NPP; ++q) {
            d[q] = 0.0f; pidx[q] = -1;
            if (t + q * TPG >= a.P) continue;
            pidx[q] = __float_as_int(sThr[t + q * TPG].w);
            d[q] = dense_pair_border(a.table, pidx[q], L, H, W, oy, ox, iy0, ix0, wgt[0], wgt[1], wgt[2], wgt[3], a.mode,
                                     a.temperature);
        }
    }
#pragma unroll
    for (int q = 0; q < NPP; ++q) ss = fmaf(d[q], d[q], ss);
    float inv = 1.0f;
    if (a.normalize) inv = 1.0f / fmaxf(sqrtf(group_sum(ss, red, g, t)), 1e-12f);
#pragma unroll
    for (int q = 0; q < NPP; ++q)
        if (pidx[q] >= 0) out[pidx[q]] = d[q] * inv;
}

// Every 64-thread group walks its keypoints (grid-stride); NBUF as in sparse_win_kernel.
// CPA: the window is fetched by the group's 64 threads as 16-byte cp.async copies instead of one TMA box (measured
// slower: 105 us against 83 us, the address arithmetic of 10 copies per thread costs more than the TMA unit saves);
// kept as a cross-check of the TMA path.
template <int GROUPS, int NBUF, bool CPA, int NPP>
__global__ void __launch_bounds__(GROUPS * TPG, 4) dense_at_kpts_kernel(const __grid_constant__ CUtensorMap tmap, DenseKpArgs a) {
    constexpr int LO = DK_LO, SPAN = DK_SPAN;   // integral rows [iy0-15, iy0+31] cover every tap of the 4 neighbours
    constexpr int WBUF = DK_ROWS * SPAN;                     // 48*52*4 bytes per window: a multiple of 128
    extern __shared__ __align__(128) float sI[];
    __shared__ float red[GROUPS * 2];
    __shared__ __align__(8) unsigned long long bars[GROUPS * 2];
    uint4* sTap = reinterpret_cast<uint4*>(sI + GROUPS * NBUF * WBUF);   // 2 x 4 window byte offsets per pair, 16 bits each
    float4* sThr = reinterpret_cast<float4*>(sTap + a.P);             // {threshold, area, float(1/area), pair index}
    const int g = threadIdx.x / TPG, t = threadIdx.x % TPG;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < GROUPS * 2; ++i) mbar_init(smem_u32(&bars[i]), 1);
        mbar_fence_init();
    }

    for (int p = threadIdx.x; p < a.P; p += GROUPS * TPG) {
        // slot p takes pair pair_of_slot(p): 32 consecutive slots are the lanes of one tap read, ordered for few bank
        // conflicts.  Window coordinates of the taps of neighbour (0,0): pixel q of the image is integral index q + MAXR,
        // the window starts at integral index iy0 - LO
        const int pair = pair_of_slot(p, a.P);
        const PairRow row = load_pair(a.table, pair);
        const int r = (int)row.r;
        const int cy1 = LO + MAXR + (int)row.oy1, cx1 = LO + MAXR + (int)row.ox1;
        const int cy2 = LO + MAXR + (int)row.oy2, cx2 = LO + MAXR + (int)row.ox2;
        auto off = [&](int y, int x) -> unsigned { return (unsigned)(y * SPAN + x) * 4u; };
        // order of bad.py:98: (y1,x1) - (y0,x1) - (y1,x0) + (y0,x0)
        static_assert((DK_ROWS + 1) * DK_SPAN * 4 < 65536, "window byte offsets (+ one neighbour row) must fit 16 bits");
        const unsigned o[8] = {off(cy1 + r + 1, cx1 + r + 1), off(cy1 - r, cx1 + r + 1), off(cy1 + r + 1, cx1 - r), off(cy1 - r, cx1 - r),
                               off(cy2 + r + 1, cx2 + r + 1), off(cy2 - r, cx2 + r + 1), off(cy2 + r + 1, cx2 - r), off(cy2 - r, cx2 - r)};
        sTap[p] = pack_taps(o);
        const float side = (float)(2 * r + 1);
        sThr[p] = make_float4(row.thr, side * side, __fdiv_rn(1.0f, side * side), __int_as_float(pair));
    }
    __syncthreads();

    const long long total = (long long)a.B * a.K;
    const long long stride = (long long)gridDim.x * GROUPS;
    float* wbuf = sI + (size_t)g * NBUF * WBUF;
    const uint32_t bar0 = smem_u32(&bars[2 * g]);
    // thread 0 of the group launches the window load of keypoint k into buffer `buf`: one TMA box, rows/columns
    // outside the integral read as zero and are never used as taps
    const int IP = ipitch(a.W, MAXR), IH = a.H + 2 * MAXR + 1;
    auto issue = [&](long long k, int buf) {
        if (a.dbg & 2) {
            if (CPA) asm volatile("cp.async.commit_group;" ::: "memory");
            else if (t == 0) mbar_arrive(bar0 + 8u * buf);
            return;
        }
        if (CPA) {
            const float ky = a.kpts[k * 2 + 0], kx = a.kpts[k * 2 + 1];
            if (ky >= 0.0f) {
                const DenseKp kp = dense_kp_pos(ky, kx, a.H, a.W);
                const int wx0 = (kp.ix0 - LO) & ~3, wy0 = kp.iy0 - LO;
                const float* img = a.I + (size_t)(k / a.K) * IH * IP;
                const uint32_t dst0 = smem_u32(wbuf + buf * WBUF);
                constexpr int C4 = SPAN / 4;                            // 16-byte pieces per window row
                for (int i = t; i < DK_ROWS * C4; i += TPG) {
                    const int row = i / C4, c4 = i - row * C4;
                    const int gy = wy0 + row, gx = wx0 + 4 * c4;
                    const bool ok = gy >= 0 && gy < IH && gx >= 0 && gx < IP;      // outside: zero fill, never used as a tap
                    const float* src = ok ? img + (size_t)gy * IP + gx : a.I;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + (uint32_t)(row * SPAN + 4 * c4) * 4u),
                                 "l"(src), "r"(ok ? 16 : 0) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        } else if (t == 0) {
            const float ky = a.kpts[k * 2 + 0], kx = a.kpts[k * 2 + 1];
            if (ky >= 0.0f) {
                const DenseKp kp = dense_kp_pos(ky, kx, a.H, a.W);
                const uint32_t bar = bar0 + 8u * buf;
                mbar_arrive_expect_tx(bar, WBUF * 4);
                tma_load_3d(smem_u32(wbuf + buf * WBUF), &tmap, bar, (kp.ix0 - LO) & ~3, kp.iy0 - LO, (int)(k / a.K));
            }
        }
    };
    long long kidx = (long long)blockIdx.x * GROUPS + g;
    if (kidx < total) issue(kidx, 0);
    uint32_t phase[2] = {0u, 0u};
    for (int n = 0; kidx < total; kidx += stride, ++n) {
        const int buf = NBUF == 2 ? (n & 1) : 0;
        if (NBUF == 2 && kidx + stride < total) issue(kidx + stride, buf ^ 1);   // that buffer was released by the barrier below
        const float ky = a.kpts[kidx * 2 + 0], kx = a.kpts[kidx * 2 + 1];
        if (ky >= 0.0f) {
            dense_kp_group<CPA, NPP>(a, kidx, ky, kx, wbuf + buf * WBUF, bar0 + 8u * buf, phase[buf], red, sTap, sThr, g, t);
            phase[buf] ^= 1u;
        } else {                                                    // shi_tomasi_bad_sinkhorn.py:143,158: masked rows are zero
            float* out = a.desc + (size_t)kidx * a.P;
            for (int p = t; p < a.P; p += TPG) out[p] = 0.0f;
        }
        group_bar(g);                                               // everyone is done with this buffer
        if (NBUF == 1 && kidx + stride < total) issue(kidx + stride, 0);
    }
}

int g_dense_window_tma = 1;    // om_debug_dense_window: bit 0: 1 = one TMA box per keypoint window (default, 83 us), 0 = 16-byte cp.async
                               // copies (105 us); bits 1, 2: diagnosis (skip the fetch / skip the pair arithmetic)

template <int GROUPS>
int launch_dense_kp(const DenseKpArgs& a, cudaStream_t st) {
    CUtensorMap tmap;
    const int IP = ipitch(a.W, MAXR);
    OM_TRY(make_tmap_3d(&tmap, true, a.I, (uint64_t)IP, (uint64_t)(a.H + 2 * MAXR + 1), (uint64_t)a.B, (uint64_t)IP, DK_SPAN,
                        DK_ROWS));
    constexpr int NBUF = 1;
    const size_t smem = (size_t)GROUPS * NBUF * DK_ROWS * DK_SPAN * sizeof(float) + (size_t)a.P * (sizeof(uint4) + sizeof(float4));
    const long long total = (long long)a.B * a.K;
    const long long nblk = (total + GROUPS - 1) / GROUPS;
    const int per_sm = (int)(228 * 1024 / (smem + 1024));          // resident CTAs per SM by shared memory (228 KB, 1 KB reserved per CTA)
    const long long resident = 148ll * (per_sm < 1 ? 1 : per_sm);
    const unsigned grid = (unsigned)(nblk < resident ? nblk : resident);
    DenseKpArgs a2 = a;
    a2.dbg = g_dense_window_tma & 6;
    const bool cpa = !(g_dense_window_tma & 1);
    auto go = [&](auto kernel) -> int {
        OM_TRY(set_smem(kernel, smem));
        kernel<<<grid, GROUPS * TPG, smem, st>>>(tmap, a2);
        return OM_OK;
    };
    if (a.P <= 4 * TPG) OM_TRY(cpa ? go(dense_at_kpts_kernel<GROUPS, NBUF, true, 4>) : go(dense_at_kpts_kernel<GROUPS, NBUF, false, 4>));
    else OM_TRY(cpa ? go(dense_at_kpts_kernel<GROUPS, NBUF, true, 8>) : go(dense_at_kpts_kernel<GROUPS, NBUF, false, 8>));
    OM_AFTER_LAUNCH();
    return OM_OK;
}

__global__ void __launch_bounds__(256) gather_kernel(const float* map, int B, int D, int H, int W, const float* kpts,
                                                     int K, int subpixel, float* out) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)B * K * D) return;
    const int dch = (int)(i % D);
    const long long bk = i / D;
    const int z = (int)(bk / K);
    const float ky = kpts[bk * 2], kx = kpts[bk * 2 + 1];
    const float* m = map + ((size_t)z * D + dch) * H * W;
    if (!subpixel) {                                                        // bad.py:258-272 (.long() truncates)
        const int y = clampi((int)ky, 0, H - 1), x = clampi((int)kx, 0, W - 1);
        out[i] = m[(size_t)y * W + x];
        return;
    }
    const float gyn = __fsub_rn(__fmul_rn(__fdiv_rn(ky, (float)(H - 1 + 1e-8)), 2.0f), 1.0f);   // bad.py:311-312
    const float gxn = __fsub_rn(__fmul_rn(__fdiv_rn(kx, (float)(W - 1 + 1e-8)), 2.0f), 1.0f);
    const float uy = fminf(fmaxf(__fmul_rn(__fadd_rn(gyn, 1.0f), (float)(H - 1) * 0.5f), 0.0f), (float)(H - 1));
    const float ux = fminf(fmaxf(__fmul_rn(__fadd_rn(gxn, 1.0f), (float)(W - 1) * 0.5f), 0.0f), (float)(W - 1));
    const float fy = floorf(uy), fx = floorf(ux);
    const float w = ux - fx, e = 1.0f - w, s = uy - fy, n = 1.0f - s;
    const int y0 = (int)fy, x0 = (int)fx, y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    out[i] = m[(size_t)y0 * W + x0] * (n * e) + m[(size_t)y0 * W + x1] * (n * w) + m[(size_t)y1 * W + x0] * (s * e) +
             m[(size_t)y1 * W + x1] * (s * w);
}

struct DenseWs {
    float* T;       // (B, H+14, W+14)   column pass of the two-pass build (float-valued images only)
    float* I;       // (B, H+15, ipitch)
    float* planes;  // (B, 8, H, W)   (dense map only)
    unsigned int* flags;    // (B+1) banded build: image has a pixel that is not a small integer; [B] = any image
    unsigned int* colsum;   // banded build: (B, bands, ipitch)
};

DenseWs carve_dense(void* ws, int B, int H, int W) {
    DenseWs d;
    char* p = (char*)ws;
    d.T = (float*)p;
    p += align_up((size_t)B * (H + 2 * MAXR) * (W + 2 * MAXR) * sizeof(float));
    d.I = (float*)p;
    p += align_up((size_t)B * (H + 2 * MAXR + 1) * ipitch(W, MAXR) * sizeof(float));
    d.planes = (float*)p;
    p += align_up((size_t)B * (MAXR + 1) * H * W * sizeof(float));
    d.flags = (unsigned int*)p;
    p += align_up((size_t)(B + 1) * sizeof(unsigned int));
    d.colsum = (unsigned int*)p;
    return d;
}

struct SparseWs {
    unsigned int* flags;   // (B+1) non-integer-pixel flag per image, [B] = any image flagged
    unsigned int* T;       // (B, H+2pad, W+2pad)   column pass of the two-pass build (very wide images only)
    unsigned int* I;       // (B, H+2pad+1, ipitch)
    unsigned int* colsum;  // banded build: (B, bands, ipitch)
};

// keypoint groups per CTA of the window kernels (10 groups per CTA, i.e. 20 windows per SM instead of 16, measured no faster)
constexpr int KPG = 4;
// window half size / integral padding per mode (see sparse_win_kernel)
// (1 + pad) % 4 == 0 so that a thread's four integral columns are four ALIGNED pixels of the image (banded build)
constexpr int HS_PLAIN = 23, PAD_PLAIN = 23, HS_ORI = 30, PAD_ORI = 31;

SparseWs carve_sparse(void* ws, int B, int H, int W, int pad) {
    SparseWs d;
    char* p = (char*)ws;
    d.flags = (unsigned int*)p;
    p += align_up((size_t)(B + 1) * sizeof(unsigned int));
    d.T = (unsigned int*)p;
    p += align_up((size_t)B * (H + 2 * pad) * (W + 2 * pad) * sizeof(unsigned int));
    d.I = (unsigned int*)p;
    p += align_up((size_t)B * (H + 2 * pad + 1) * ipitch(W, pad) * sizeof(unsigned int));
    d.colsum = (unsigned int*)p;
    return d;
}

int check_table(const float* table, int P) {
    if (table == nullptr) return OM_ERR_NULL;
    if (P <= 0 || P > 512) return OM_ERR_PARAM;
    return OM_OK;
}

// float32 integral of the dense path.  Integer-valued pixels (every uint8 image; float images are checked): banded exact
// build, RN(exact) == the reference's double-accumulated cumsums.  A float image with a non-integer (or huge) pixel raises
// the batch flag and the two-pass double-accumulating build, launched behind and gated on that flag, redoes the batch.
int build_integral(const void* image, int image_u8, int B, int H, int W, const DenseWs& d, cudaStream_t st) {
    const double cells = (double)(H + 2 * MAXR) * (double)(W + 2 * MAXR);
    const float maxv = (float)fmin(65535.0, floor(4294967295.0 / cells));     // every partial sum stays below 2^32
    OM_CUDA(cudaMemsetAsync(d.flags, 0, (size_t)(B + 1) * sizeof(unsigned int), st));
    int rc;
    if (image_u8) rc = build_integral_banded<unsigned char, float>((const unsigned char*)image, B, H, W, MAXR, maxv, d.colsum, d.I, d.flags, st);
    else rc = build_integral_banded<float, float>((const float*)image, B, H, W, MAXR, maxv, d.colsum, d.I, d.flags, st);
    if (rc == OM_ERR_LIMIT) {                                 // too wide for one CTA per band
        if (image_u8) return OM_ERR_LIMIT;
        return build_prefix<false>((const float*)image, B, H, W, MAXR, d.T, d.I, nullptr, st);
    }
    OM_TRY(rc);
    if (image_u8) return OM_OK;
    return build_prefix<false>((const float*)image, B, H, W, MAXR, d.T, d.I, nullptr, st, d.flags + B);
}

}  // namespace

size_t sparse_bad_workspace_bytes(int B, int H, int W, int theta_mode) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    const int pad = theta_mode == OM_THETA_NONE ? PAD_PLAIN : PAD_ORI;
    return align_up((size_t)(B + 1) * sizeof(unsigned int)) +
           align_up((size_t)B * (H + 2 * pad) * (W + 2 * pad) * sizeof(unsigned int)) +
           align_up((size_t)B * (H + 2 * pad + 1) * ipitch(W, pad) * sizeof(unsigned int)) + band_colsum_bytes(B, H, W, pad);
}

int sparse_bad_launch(const void* image, int image_u8, int B, int H, int W, const float* kpts, int K, const float* pair_table,
                      int P, int desc_mode, float temperature, int normalize, int sampling_mode, int theta_mode,
                      const float* orientation, const float* moment_kernels, int patch_size, float* desc,
                      void* ws, size_t ws_bytes, cudaStream_t st, int phase) {
    // phase 0: everything; 1: only the integral image (needs no keypoints); 2: only the descriptors (integral built)
    if (image == nullptr || kpts == nullptr || desc == nullptr) return OM_ERR_NULL;
    if (B <= 0 || H <= 1 || W <= 1 || K <= 0) return OM_ERR_SHAPE;
    OM_TRY(check_table(pair_table, P));
    if (desc_mode < OM_DESC_RAW || desc_mode > OM_DESC_HARD) return OM_ERR_PARAM;
    if (sampling_mode != OM_SAMPLE_NEAREST && sampling_mode != OM_SAMPLE_BILINEAR) return OM_ERR_PARAM;
    if (theta_mode < OM_THETA_NONE || theta_mode > OM_THETA_MOMENTS) return OM_ERR_PARAM;
    if (theta_mode == OM_THETA_MAP && orientation == nullptr) return OM_ERR_NULL;
    if (theta_mode == OM_THETA_MOMENTS) {
        if (moment_kernels == nullptr) return OM_ERR_NULL;
        if (patch_size < 1 || patch_size % 2 == 0 || patch_size > 31) return OM_ERR_PARAM;
    }
    if (ws == nullptr || ws_bytes < sparse_bad_workspace_bytes(B, H, W, theta_mode)) return OM_ERR_WORKSPACE;
    const bool oriented = theta_mode != OM_THETA_NONE;
    const int pad = oriented ? PAD_ORI : PAD_PLAIN;
    const SparseWs w = carve_sparse(ws, B, H, W, pad);
    // 16-bit integral (modulo 2^16) whenever the banded build applies (a pure function of the sizes: phases 1 and 2 of a call
    // agree): pixels are then held to [0, 255] by the integer-image test (anything else raises the image's flag and goes to
    // the general kernel), so that a 15 x 15 box sum stays below 2^16
    const bool i16 = g_sparse_i16 && band_colsum_pitch(W, pad) <= IB_PX * 1024 && B <= 65535;
    if (phase != 2) {
        OM_CUDA(cudaMemsetAsync(w.flags, 0, (size_t)(B + 1) * sizeof(unsigned int), st));
        int rc;
        if (i16) {
            unsigned short* I16 = reinterpret_cast<unsigned short*>(w.I);
            if (image_u8) rc = build_integral_banded<unsigned char, unsigned short>((const unsigned char*)image, B, H, W, pad, 255.0f, w.colsum, I16, w.flags, st);
            else rc = build_integral_banded<float, unsigned short>((const float*)image, B, H, W, pad, 255.0f, w.colsum, I16, w.flags, st);
        } else {
            if (image_u8) rc = build_integral_banded<unsigned char, unsigned int>((const unsigned char*)image, B, H, W, pad, 65535.0f, w.colsum, w.I, w.flags, st);
            else rc = build_integral_banded<float, unsigned int>((const float*)image, B, H, W, pad, 65535.0f, w.colsum, w.I, w.flags, st);
            if (rc == OM_ERR_LIMIT && !image_u8) rc = build_prefix<true>((const float*)image, B, H, W, pad, w.T, w.I, w.flags, st);
        }
        OM_TRY(rc);
    }
    if (phase == 1) return OM_OK;
    SparseArgs a{};
    a.image = image; a.image_u8 = image_u8; a.B = B; a.H = H; a.W = W; a.kpts = kpts; a.K = K; a.table = pair_table; a.P = P;
    a.mode = desc_mode; a.temperature = temperature; a.normalize = normalize;
    a.bilinear = sampling_mode == OM_SAMPLE_BILINEAR; a.theta_mode = theta_mode; a.orientation = orientation;
    a.moments = moment_kernels; a.patch_size = patch_size; a.desc = desc;
    a.sy = (float)(2.0 / ((double)(H - 1) + 1e-8));
    a.sx = (float)(2.0 / ((double)(W - 1) + 1e-8));
    a.flags = w.flags; a.pad = pad;
    const bool bil = a.bilinear != 0;
    // integer-valued images: window kernel on the exact integral; every other image: the general kernel
    using U32 = unsigned int;
    using U16 = unsigned short;
    if (!oriented) {
        if (i16) OM_TRY((bil ? launch_sparse_win<HS_PLAIN, KPG, false, true, U16>(a, w.I, st) : launch_sparse_win<HS_PLAIN, KPG, false, false, U16>(a, w.I, st)));
        else OM_TRY((bil ? launch_sparse_win<HS_PLAIN, KPG, false, true, U32>(a, w.I, st) : launch_sparse_win<HS_PLAIN, KPG, false, false, U32>(a, w.I, st)));
        return bil ? launch_sparse<HS_PLAIN, 4, false, true>(a, st) : launch_sparse<HS_PLAIN, 4, false, false>(a, st);
    }
    if (i16) OM_TRY((bil ? launch_sparse_win<HS_ORI, 4, true, true, U16>(a, w.I, st) : launch_sparse_win<HS_ORI, 4, true, false, U16>(a, w.I, st)));
    else OM_TRY((bil ? launch_sparse_win<HS_ORI, 4, true, true, U32>(a, w.I, st) : launch_sparse_win<HS_ORI, 4, true, false, U32>(a, w.I, st)));
    return bil ? launch_sparse<HS_ORI, 3, true, true>(a, st) : launch_sparse<HS_ORI, 3, true, false>(a, st);
}

size_t dense_bad_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return align_up((size_t)B * (H + 2 * MAXR) * (W + 2 * MAXR) * sizeof(float)) +
           align_up((size_t)B * (H + 2 * MAXR + 1) * ipitch(W, MAXR) * sizeof(float)) +
           align_up((size_t)B * (MAXR + 1) * H * W * sizeof(float)) + align_up((size_t)(B + 1) * sizeof(unsigned int)) +
           band_colsum_bytes(B, H, W, MAXR);
}

int dense_bad_at_kpts_launch(const void* image, int image_u8, int B, int H, int W, const float* kpts, int K,
                             const float* pair_table, int P, int desc_mode, float temperature, int normalize,
                             float* desc, void* ws, size_t ws_bytes, cudaStream_t st, int phase) {
    if (image == nullptr || kpts == nullptr || desc == nullptr) return OM_ERR_NULL;
    if (B <= 0 || H <= 1 || W <= 1 || K <= 0) return OM_ERR_SHAPE;
    OM_TRY(check_table(pair_table, P));
    if (desc_mode < OM_DESC_RAW || desc_mode > OM_DESC_HARD) return OM_ERR_PARAM;
    if (ws == nullptr || ws_bytes < dense_bad_workspace_bytes(B, H, W)) return OM_ERR_WORKSPACE;
    const DenseWs d = carve_dense(ws, B, H, W);
    if (phase != 2) OM_TRY(build_integral(image, image_u8, B, H, W, d, st));
    if (phase == 1) return OM_OK;
    DenseKpArgs a{};
    a.I = d.I; a.B = B; a.H = H; a.W = W; a.kpts = kpts; a.K = K; a.table = pair_table; a.P = P; a.mode = desc_mode;
    a.temperature = temperature; a.normalize = normalize; a.desc = desc;
    return launch_dense_kp<KPG>(a, st);
}

}  // namespace om

using namespace om;

extern "C" void om_debug_dense_window(int tma) { g_dense_window_tma = tma; }
extern "C" void om_debug_band_rows(int rows) {
    if (rows == -16 || rows == -32) { g_sparse_i16 = rows == -16; return; }     // sparse path: 16-bit / 32-bit integral
    g_band_rows = rows >= 8 && rows <= 256 ? rows / 2 * 2 : 32;
}

extern "C" int om_angle_map_f32(const float* image, int B, int H, int W, const float* moment_kernels, int patch_size,
                                float* angle_map, void* stream) {
    OM_ON_DEVICE_OF(image);
    if (image == nullptr || moment_kernels == nullptr || angle_map == nullptr) return OM_ERR_NULL;
    if (B <= 0 || H <= 0 || W <= 0) return OM_ERR_SHAPE;
    if (patch_size < 1 || patch_size % 2 == 0 || patch_size > 31) return OM_ERR_PARAM;
    const int half = patch_size / 2;
    const size_t smem = ((size_t)2 * patch_size * patch_size + (size_t)(ATH + 2 * half) * ((ATW + 2 * half) | 1)) * sizeof(float);
    OM_TRY(set_smem(angle_map_kernel, smem));
    angle_map_kernel<<<dim3((W + ATW - 1) / ATW, (H + ATH - 1) / ATH, B), 256, smem, (cudaStream_t)stream>>>(
        image, H, W, moment_kernels, patch_size, angle_map);
    OM_AFTER_LAUNCH();
    return OM_OK;
}

extern "C" int om_sparse_bad_f32(const float* image, int B, int H, int W, const float* kpts, int K,
                                 const float* pair_table, int P, int desc_mode, float temperature, int normalize,
                                 int sampling_mode, int theta_mode, const float* orientation,
                                 const float* moment_kernels, int patch_size, float* desc, void* ws, size_t ws_bytes,
                                 void* stream) {
    OM_ON_DEVICE_OF(image);
    return sparse_bad_launch(image, 0, B, H, W, kpts, K, pair_table, P, desc_mode, temperature, normalize, sampling_mode,
                             theta_mode, orientation, moment_kernels, patch_size, desc, ws, ws_bytes,
                             (cudaStream_t)stream);
}

extern "C" size_t om_sparse_bad_workspace_bytes(int B, int H, int W, int theta_mode) {
    return sparse_bad_workspace_bytes(B, H, W, theta_mode);
}

extern "C" size_t om_dense_bad_workspace_bytes(int B, int H, int W) { return dense_bad_workspace_bytes(B, H, W); }

extern "C" int om_dense_bad_f32(const float* image, int B, int H, int W, const float* pair_table, int P, int desc_mode,
                                float temperature, float* desc_map, void* ws, size_t ws_bytes, void* stream) {
    OM_ON_DEVICE_OF(image);
    if (image == nullptr || desc_map == nullptr) return OM_ERR_NULL;
    if (B <= 0 || H <= 0 || W <= 0 || B > 65535 || H > 65535) return OM_ERR_SHAPE;
    OM_TRY(check_table(pair_table, P));
    if (desc_mode < OM_DESC_RAW || desc_mode > OM_DESC_HARD) return OM_ERR_PARAM;
    if (ws == nullptr || ws_bytes < dense_bad_workspace_bytes(B, H, W)) return OM_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const DenseWs d = carve_dense(ws, B, H, W);
    OM_TRY(build_integral(image, 0, B, H, W, d, st));
    const dim3 grid((W + 127) / 128, H, B);
    box_planes_kernel<<<grid, 128, 0, st>>>(d.I, H, W, d.planes);
    OM_AFTER_LAUNCH();
    dense_map_kernel<<<grid, 128, 0, st>>>(d.planes, H, W, pair_table, P, desc_mode, temperature, desc_map);
    OM_AFTER_LAUNCH();
    return OM_OK;
}

extern "C" int om_dense_bad_at_kpts_f32(const float* image, int B, int H, int W, const float* kpts, int K,
                                        const float* pair_table, int P, int desc_mode, float temperature,
                                        int normalize, float* desc, void* ws, size_t ws_bytes, void* stream) {
    OM_ON_DEVICE_OF(image);
    return dense_bad_at_kpts_launch(image, 0, B, H, W, kpts, K, pair_table, P, desc_mode, temperature, normalize, desc, ws,
                                    ws_bytes, (cudaStream_t)stream);
}

extern "C" int om_gather_descriptors_f32(const float* desc_map, int B, int D, int H, int W, const float* kpts, int K,
                                         int subpixel, float* desc, void* stream) {
    OM_ON_DEVICE_OF(desc_map);
    if (desc_map == nullptr || kpts == nullptr || desc == nullptr) return OM_ERR_NULL;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || K <= 0) return OM_ERR_SHAPE;
    const long long n = (long long)B * K * D;
    gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(desc_map, B, D, H, W, kpts, K, subpixel,
                                                                                desc);
    OM_AFTER_LAUNCH();
    return OM_OK;
}

// stage 0: the two integral-image kernels; stage 1: the keypoint kernel only (expects stage 0 ran on ws)
extern "C" int om_debug_dense_stage(const float* image, int B, int H, int W, const float* kpts, int K,
                                    const float* pair_table, int P, int desc_mode, float temperature, int normalize,
                                    float* desc, void* ws, size_t ws_bytes, void* stream, int stage) {
    OM_ON_DEVICE_OF(image);
    if (image == nullptr || kpts == nullptr || desc == nullptr) return OM_ERR_NULL;
    if (B <= 0 || H <= 1 || W <= 1 || K <= 0) return OM_ERR_SHAPE;
    OM_TRY(check_table(pair_table, P));
    if (ws == nullptr || ws_bytes < dense_bad_workspace_bytes(B, H, W)) return OM_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const DenseWs d = carve_dense(ws, B, H, W);
    if (stage == 0) return build_integral(image, 0, B, H, W, d, st);
    DenseKpArgs a{};
    a.I = d.I; a.B = B; a.H = H; a.W = W; a.kpts = kpts; a.K = K; a.table = pair_table; a.P = P; a.mode = desc_mode;
    a.temperature = temperature; a.normalize = normalize; a.desc = desc;
    return launch_dense_kp<KPG>(a, st);
}
