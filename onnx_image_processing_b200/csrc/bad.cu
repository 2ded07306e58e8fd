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

#include "common.cuh"

namespace om {

namespace {

constexpr int TPG = 64;          // threads per keypoint group
constexpr int MAXR = 7;          // largest box radius in the tables

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

__device__ __forceinline__ float finish_value(float diff, float thr, int mode, float temperature) {
    const float centered = __fsub_rn(diff, thr);                       // bad.py:212 / :559
    if (mode == OM_DESC_RAW) return centered;
    if (mode == OM_DESC_SOFT) {                                        // sigmoid(-centered * T)
        const float x = __fmul_rn(-centered, temperature);
        return 1.0f / (1.0f + expf(-x));
    }
    return centered <= 0.0f ? 1.0f : 0.0f;
}

// sum of `v` over the TPG threads of one keypoint group (2 warps); red is per-CTA scratch [groups][2]
__device__ __forceinline__ float group_sum(float v, float* red, int g, int t) {
    v = warp_sum(v);
    __syncthreads();
    if ((t & 31) == 0) red[g * 2 + (t >> 5)] = v;
    __syncthreads();
    return red[g * 2] + red[g * 2 + 1];
}

struct SparseArgs {
    const float* image;
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
};

// HS: patch half size; GROUPS keypoints per CTA
template <int HS, int GROUPS>
__global__ void __launch_bounds__(GROUPS * TPG) sparse_bad_kernel(SparseArgs a) {
    constexpr int S = 2 * HS + 1;       // patch side
    constexpr int PD = S + 2;           // odd pitch (in doubles) of the (S+1)x(S+1) integral
    extern __shared__ __align__(16) double sD[];
    __shared__ float red[GROUPS * 2];
    __shared__ float sTheta[GROUPS];

    const int g = threadIdx.x / TPG, t = threadIdx.x % TPG;
    const long long total = (long long)a.B * a.K;
    long long kidx = (long long)blockIdx.x * GROUPS + g;
    const bool live = kidx < total;
    if (!live) kidx = total - 1;
    const int z = (int)(kidx / a.K);
    const int H = a.H, W = a.W;
    const float* img = a.image + (size_t)z * H * W;
    double* D = sD + (size_t)g * (S + 1) * PD;

    const float ky = a.kpts[kidx * 2 + 0], kx = a.kpts[kidx * 2 + 1];
    const float valid = ky >= 0.0f ? 1.0f : 0.0f;                          // bad.py:461
    const float yc = fminf(fmaxf(ky, 0.0f), (float)(H - 1));               // bad.py:464-465
    const float xc = fminf(fmaxf(kx, 0.0f), (float)(W - 1));
    const int iy0 = (int)nearbyintf(yc), ix0 = (int)nearbyintf(xc);
    const int py0 = iy0 - HS, px0 = ix0 - HS;                               // global coords of patch (0,0)

    // replicate-clamped patch (bad.py:474-478 pads the image, grid_sample clamps the centre).
    // Thread t owns integral column t (patch column t-1): its clamped global column is fixed, rows
    // are independent loads.  Row 0 / column 0 of the integral are zero.
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
    __syncthreads();

    // orientation of this keypoint
    float ct = 1.0f, st = 0.0f;
    if (a.theta_mode != OM_THETA_NONE) {
        // nearest sample of the orientation map at the keypoint (bad.py:490-499)
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
        }
        if (t == 0) sTheta[g] = theta;
        __syncthreads();
        theta = sTheta[g];
        ct = cosf(theta);                                                   // bad.py:501-502
        st = sinf(theta);
    }
    __syncthreads();

    // exact integral image of the patch: column prefix, then row prefix
    if (t < S) {
        double acc = 0.0;
        double* p = D + PD + (t + 1);
        for (int y = 0; y < S; ++y, p += PD) { acc += *p; *p = acc; }
    }
    __syncthreads();
    if (t < S) {
        double acc = 0.0;
        double* p = D + (t + 1) * PD + 1;
        for (int x = 0; x < S; ++x) { acc += p[x]; p[x] = acc; }
    }
    __syncthreads();

    const float hy = (float)(H - 1) * 0.5f, hx = (float)(W - 1) * 0.5f;
    const float my = (float)(H - 1), mx = (float)(W - 1);

    auto box_mean = [&](int cy, int cx, int r) -> float {
        // cy,cx: integer sample centre in image coords; box of radius r on the replicate-padded image
        const int pyc = clampi(cy - py0, r, S - 1 - r), pxc = clampi(cx - px0, r, S - 1 - r);
        const int y0 = pyc - r, y1 = pyc + r + 1, x0 = pxc - r, x1 = pxc + r + 1;
        const double s = D[y1 * PD + x1] - D[y0 * PD + x1] - D[y1 * PD + x0] + D[y0 * PD + x0];
        const double side = (double)(2 * r + 1);
        return (float)(s / (side * side));
    };
    auto sample = [&](float oy, float ox, int r) -> float {
        float py, px;
        if (a.theta_mode != OM_THETA_NONE) {                                // bad.py:504-517
            const float dy = __fadd_rn(__fmul_rn(ox, st), __fmul_rn(oy, ct));
            const float dx = __fsub_rn(__fmul_rn(ox, ct), __fmul_rn(oy, st));
            py = __fadd_rn(yc, dy);
            px = __fadd_rn(xc, dx);
        } else {                                                            // bad.py:518-525
            py = __fadd_rn(yc, oy);
            px = __fadd_rn(xc, ox);
        }
        const float uy = sample_coord(py, a.sy, hy, my), ux = sample_coord(px, a.sx, hx, mx);
        if (!a.bilinear) return box_mean((int)nearbyintf(uy), (int)nearbyintf(ux), r);   // half-to-even
        const float fy = floorf(uy), fx = floorf(ux);
        const float w = ux - fx, e = 1.0f - w, s = uy - fy, n = 1.0f - s;   // ATen bilinear weights
        const int y_n = (int)fy, x_w = (int)fx;
        const int y_s = min(y_n + 1, H - 1), x_e = min(x_w + 1, W - 1);     // weight is 0 where this clamps
        const float nw = box_mean(y_n, x_w, r), ne = box_mean(y_n, x_e, r);
        const float sw = box_mean(y_s, x_w, r), se = box_mean(y_s, x_e, r);
        return nw * (n * e) + ne * (n * w) + sw * (s * e) + se * (s * w);
    };

    constexpr int MAXPP = 8;   // pairs per thread: P <= 512
    float d[MAXPP];
    float ss = 0.0f;
#pragma unroll
    for (int q = 0; q < MAXPP; ++q) {
        const int p = t + q * TPG;
        d[q] = 0.0f;
        if (p < a.P) {
            const PairRow row = load_pair(a.table, p);
            const int r = (int)row.r;
            const float diff = __fsub_rn(sample(row.oy1, row.ox1, r), sample(row.oy2, row.ox2, r));   // bad.py:557
            d[q] = __fmul_rn(finish_value(diff, row.thr, a.mode, a.temperature), valid);            // bad.py:570
            ss = fmaf(d[q], d[q], ss);
        }
    }
    float inv = 1.0f;
    if (a.normalize) {                                                      // F.normalize, bad.py:573-574
        const float nrm = sqrtf(group_sum(ss, red, g, t));
        inv = 1.0f / fmaxf(nrm, 1e-12f);
    }
    if (live) {
        float* out = a.desc + (size_t)kidx * a.P;
#pragma unroll
        for (int q = 0; q < MAXPP; ++q) {
            const int p = t + q * TPG;
            if (p < a.P) out[p] = a.normalize ? d[q] * inv : d[q];
        }
    }
}

template <int HS, int GROUPS>
int launch_sparse(const SparseArgs& a, cudaStream_t st) {
    constexpr int S = 2 * HS + 1;
    constexpr size_t smem = (size_t)GROUPS * (S + 1) * (S + 2) * sizeof(double);
    OM_TRY(set_smem(sparse_bad_kernel<HS, GROUPS>, smem));
    const long long total = (long long)a.B * a.K;
    const unsigned grid = (unsigned)((total + GROUPS - 1) / GROUPS);
    sparse_bad_kernel<HS, GROUPS><<<grid, GROUPS * TPG, smem, st>>>(a);
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
// pass 1: cumsum along H of the replicate-padded image, double accumulator, rounded to f32 (bad.py:70-71)
__global__ void __launch_bounds__(128) integral_cols_kernel(const float* image, int H, int W, float* T) {
    const int Wp = W + 2 * MAXR, Hp = H + 2 * MAXR;
    const int xx = blockIdx.x * 128 + threadIdx.x;
    if (xx >= Wp) return;
    const int z = blockIdx.y;
    const float* col = image + (size_t)z * H * W + clampi(xx - MAXR, 0, W - 1);
    float* dst = T + (size_t)z * Hp * Wp + xx;
    double acc = 0.0;
    for (int yy = 0; yy < Hp; ++yy) {
        acc += (double)__ldg(col + (size_t)clampi(yy - MAXR, 0, H - 1) * W);
        dst[(size_t)yy * Wp] = (float)acc;
    }
}

// pass 2: cumsum along W of pass 1 (double accumulator, rounded to f32), stored with the zero
// row/column of bad.py:72 -> I is (H+15) x (W+15)
__global__ void __launch_bounds__(256) integral_rows_kernel(const float* T, int H, int W, float* I) {
    const int Wp = W + 2 * MAXR, Hp = H + 2 * MAXR;
    const int z = blockIdx.y;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);      // row of I, 0 .. Hp
    const int lane = threadIdx.x & 31;
    if (row > Hp) return;
    float* dst = I + (size_t)z * (Hp + 1) * (Wp + 1) + (size_t)row * (Wp + 1);
    if (row == 0) {
        for (int x = lane; x <= Wp; x += 32) dst[x] = 0.0f;
        return;
    }
    const float* src = T + (size_t)z * Hp * Wp + (size_t)(row - 1) * Wp;
    if (lane == 0) dst[0] = 0.0f;
    double carry = 0.0;
    for (int x0 = 0; x0 < Wp; x0 += 32) {
        const int x = x0 + lane;
        double v = x < Wp ? (double)src[x] : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += n;
        }
        v += carry;
        if (x < Wp) dst[x + 1] = (float)v;
        carry = __shfl_sync(0xffffffffu, v, 31);
    }
}

// box mean of the reference's dense path at padded-integral centre (cy,cx) (already +MAXR):
// bad.py:87-99, fp32 taps in the reference's order, fp32 divide by the area
__device__ __forceinline__ float dense_box_mean(const float* I, int pitch, int cy, int cx, int r) {
    const int y0 = cy - r, x0 = cx - r, y1 = cy + r + 1, x1 = cx + r + 1;
    const float s = __fadd_rn(__fsub_rn(__fsub_rn(I[y1 * pitch + x1], I[y0 * pitch + x1]), I[y1 * pitch + x0]),
                              I[y0 * pitch + x0]);
    const float side = (float)(2 * r + 1);
    return __fdiv_rn(s, side * side);
}

// planes[r][y][x] = box mean of radius r centred on image pixel (y,x), r = 0..7
__global__ void __launch_bounds__(128) box_planes_kernel(const float* I, int H, int W, float* planes) {
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int y = blockIdx.y;
    const int z = blockIdx.z;
    if (x >= W) return;
    const int pitch = W + 2 * MAXR + 1;
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
};

// Dense-path descriptors at the K keypoints only: bilinear blend of the dense map's values at the
// four neighbouring pixels (bad.py:277-333), mask, L2 normalise (shi_tomasi_bad_sinkhorn.py:143-158, :213-214)
template <int GROUPS>
__global__ void __launch_bounds__(GROUPS * TPG) dense_at_kpts_kernel(DenseKpArgs a) {
    constexpr int LO = 15, SPAN = 47;       // integral rows [iy0-15, iy0+31] cover every tap of the 4 neighbours
    extern __shared__ float sI[];
    __shared__ float red[GROUPS * 2];
    const int g = threadIdx.x / TPG, t = threadIdx.x % TPG;
    const long long total = (long long)a.B * a.K;
    long long kidx = (long long)blockIdx.x * GROUPS + g;
    const bool live = kidx < total;
    if (!live) kidx = total - 1;
    const int z = (int)(kidx / a.K);
    const int H = a.H, W = a.W;
    const int Hi = H + 2 * MAXR + 1, Wi = W + 2 * MAXR + 1;
    const float* Iz = a.I + (size_t)z * Hi * Wi;
    float* L = sI + (size_t)g * SPAN * SPAN;

    const float ky = a.kpts[kidx * 2 + 0], kx = a.kpts[kidx * 2 + 1];
    const float valid = ky >= 0.0f ? 1.0f : 0.0f;
    const float yc = fminf(fmaxf(ky, 0.0f), (float)(H - 1));
    const float xc = fminf(fmaxf(kx, 0.0f), (float)(W - 1));
    // bad.py:311-312: kp / (dim-1+1e-8) * 2 - 1, then ATen unnormalise + clip
    const float gyn = __fsub_rn(__fmul_rn(__fdiv_rn(yc, (float)(H - 1 + 1e-8)), 2.0f), 1.0f);
    const float gxn = __fsub_rn(__fmul_rn(__fdiv_rn(xc, (float)(W - 1 + 1e-8)), 2.0f), 1.0f);
    const float uy = fminf(fmaxf(__fmul_rn(__fadd_rn(gyn, 1.0f), (float)(H - 1) * 0.5f), 0.0f), (float)(H - 1));
    const float ux = fminf(fmaxf(__fmul_rn(__fadd_rn(gxn, 1.0f), (float)(W - 1) * 0.5f), 0.0f), (float)(W - 1));
    const float fy = floorf(uy), fx = floorf(ux);
    const float w = ux - fx, e = 1.0f - w, s = uy - fy, n = 1.0f - s;
    const int iy0 = (int)fy, ix0 = (int)fx;
    const int ys = min(iy0 + 1, H - 1), xe = min(ix0 + 1, W - 1);
    const int oy = iy0 - LO, ox = ix0 - LO;     // integral coords of L(0,0)

    static_assert(SPAN <= TPG, "one thread per window column");
    if (t < SPAN) {
        const float* src = Iz + clampi(ox + t, 0, Wi - 1);
#pragma unroll 8
        for (int dy = 0; dy < SPAN; ++dy) L[dy * SPAN + t] = __ldg(src + (size_t)clampi(oy + dy, 0, Hi - 1) * Wi);
    }
    __syncthreads();

    auto value_at = [&](int py, int px, const PairRow& row, int r) -> float {
        const int c1y = clampi(py + (int)row.oy1, 0, H - 1) + MAXR - oy, c1x = clampi(px + (int)row.ox1, 0, W - 1) + MAXR - ox;
        const int c2y = clampi(py + (int)row.oy2, 0, H - 1) + MAXR - oy, c2x = clampi(px + (int)row.ox2, 0, W - 1) + MAXR - ox;
        const float diff = __fsub_rn(dense_box_mean(L, SPAN, c1y, c1x, r), dense_box_mean(L, SPAN, c2y, c2x, r));
        return finish_value(diff, row.thr, a.mode, a.temperature);
    };

    constexpr int MAXPP = 8;
    float d[MAXPP];
    float ss = 0.0f;
#pragma unroll
    for (int q = 0; q < MAXPP; ++q) {
        const int p = t + q * TPG;
        d[q] = 0.0f;
        if (p < a.P) {
            const PairRow row = load_pair(a.table, p);
            const int r = (int)row.r;
            const float nw = value_at(iy0, ix0, row, r), ne = value_at(iy0, xe, row, r);
            const float sw = value_at(ys, ix0, row, r), se = value_at(ys, xe, row, r);
            const float v = nw * (n * e) + ne * (n * w) + sw * (s * e) + se * (s * w);
            d[q] = __fmul_rn(v, valid);
            ss = fmaf(d[q], d[q], ss);
        }
    }
    float inv = 1.0f;
    if (a.normalize) {
        float v = warp_sum(ss);
        __syncthreads();
        if ((t & 31) == 0) red[g * 2 + (t >> 5)] = v;
        __syncthreads();
        inv = 1.0f / fmaxf(sqrtf(red[g * 2] + red[g * 2 + 1]), 1e-12f);
    }
    if (live) {
        float* out = a.desc + (size_t)kidx * a.P;
#pragma unroll
        for (int q = 0; q < MAXPP; ++q) {
            const int p = t + q * TPG;
            if (p < a.P) out[p] = a.normalize ? d[q] * inv : d[q];
        }
    }
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
    float* T;       // (B, H+14, W+14)
    float* I;       // (B, H+15, W+15)
    float* planes;  // (B, 8, H, W)   (dense map only)
};

DenseWs carve_dense(void* ws, int B, int H, int W) {
    DenseWs d;
    char* p = (char*)ws;
    d.T = (float*)p;
    p += align_up((size_t)B * (H + 2 * MAXR) * (W + 2 * MAXR) * sizeof(float));
    d.I = (float*)p;
    p += align_up((size_t)B * (H + 2 * MAXR + 1) * (W + 2 * MAXR + 1) * sizeof(float));
    d.planes = (float*)p;
    return d;
}

int check_table(const float* table, int P) {
    if (table == nullptr) return OM_ERR_NULL;
    if (P <= 0 || P > 512) return OM_ERR_PARAM;
    return OM_OK;
}

int build_integral(const float* image, int B, int H, int W, const DenseWs& d, cudaStream_t st) {
    const int Wp = W + 2 * MAXR, Hp = H + 2 * MAXR;
    integral_cols_kernel<<<dim3((Wp + 127) / 128, B), 128, 0, st>>>(image, H, W, d.T);
    OM_AFTER_LAUNCH();
    integral_rows_kernel<<<dim3((Hp + 1 + 7) / 8, B), 256, 0, st>>>(d.T, H, W, d.I);
    OM_AFTER_LAUNCH();
    return OM_OK;
}

}  // namespace

int sparse_bad_launch(const float* image, int B, int H, int W, const float* kpts, int K, const float* pair_table,
                      int P, int desc_mode, float temperature, int normalize, int sampling_mode, int theta_mode,
                      const float* orientation, const float* moment_kernels, int patch_size, float* desc,
                      cudaStream_t st) {
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
    SparseArgs a{};
    a.image = image; a.B = B; a.H = H; a.W = W; a.kpts = kpts; a.K = K; a.table = pair_table; a.P = P;
    a.mode = desc_mode; a.temperature = temperature; a.normalize = normalize;
    a.bilinear = sampling_mode == OM_SAMPLE_BILINEAR; a.theta_mode = theta_mode; a.orientation = orientation;
    a.moments = moment_kernels; a.patch_size = patch_size; a.desc = desc;
    a.sy = (float)(2.0 / ((double)(H - 1) + 1e-8));
    a.sx = (float)(2.0 / ((double)(W - 1) + 1e-8));
    if (theta_mode == OM_THETA_NONE) return launch_sparse<24, 4>(a, st);
    return launch_sparse<30, 3>(a, st);
}

size_t dense_bad_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return align_up((size_t)B * (H + 2 * MAXR) * (W + 2 * MAXR) * sizeof(float)) +
           align_up((size_t)B * (H + 2 * MAXR + 1) * (W + 2 * MAXR + 1) * sizeof(float)) +
           align_up((size_t)B * (MAXR + 1) * H * W * sizeof(float));
}

int dense_bad_at_kpts_launch(const float* image, int B, int H, int W, const float* kpts, int K,
                             const float* pair_table, int P, int desc_mode, float temperature, int normalize,
                             float* desc, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (image == nullptr || kpts == nullptr || desc == nullptr) return OM_ERR_NULL;
    if (B <= 0 || H <= 1 || W <= 1 || K <= 0) return OM_ERR_SHAPE;
    OM_TRY(check_table(pair_table, P));
    if (desc_mode < OM_DESC_RAW || desc_mode > OM_DESC_HARD) return OM_ERR_PARAM;
    if (ws == nullptr || ws_bytes < dense_bad_workspace_bytes(B, H, W)) return OM_ERR_WORKSPACE;
    const DenseWs d = carve_dense(ws, B, H, W);
    OM_TRY(build_integral(image, B, H, W, d, st));
    DenseKpArgs a{};
    a.I = d.I; a.B = B; a.H = H; a.W = W; a.kpts = kpts; a.K = K; a.table = pair_table; a.P = P; a.mode = desc_mode;
    a.temperature = temperature; a.normalize = normalize; a.desc = desc;
    constexpr int GROUPS = 4;
    constexpr size_t smem = (size_t)GROUPS * 47 * 47 * sizeof(float);
    const long long total = (long long)B * K;
    dense_at_kpts_kernel<GROUPS><<<(unsigned)((total + GROUPS - 1) / GROUPS), GROUPS * TPG, smem, st>>>(a);
    OM_AFTER_LAUNCH();
    return OM_OK;
}

}  // namespace om

using namespace om;

extern "C" int om_angle_map_f32(const float* image, int B, int H, int W, const float* moment_kernels, int patch_size,
                                float* angle_map, void* stream) {
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
                                 const float* moment_kernels, int patch_size, float* desc, void* stream) {
    return sparse_bad_launch(image, B, H, W, kpts, K, pair_table, P, desc_mode, temperature, normalize, sampling_mode,
                             theta_mode, orientation, moment_kernels, patch_size, desc, (cudaStream_t)stream);
}

extern "C" size_t om_dense_bad_workspace_bytes(int B, int H, int W) { return dense_bad_workspace_bytes(B, H, W); }

extern "C" int om_dense_bad_f32(const float* image, int B, int H, int W, const float* pair_table, int P, int desc_mode,
                                float temperature, float* desc_map, void* ws, size_t ws_bytes, void* stream) {
    if (image == nullptr || desc_map == nullptr) return OM_ERR_NULL;
    if (B <= 0 || H <= 0 || W <= 0 || B > 65535 || H > 65535) return OM_ERR_SHAPE;
    OM_TRY(check_table(pair_table, P));
    if (desc_mode < OM_DESC_RAW || desc_mode > OM_DESC_HARD) return OM_ERR_PARAM;
    if (ws == nullptr || ws_bytes < dense_bad_workspace_bytes(B, H, W)) return OM_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const DenseWs d = carve_dense(ws, B, H, W);
    OM_TRY(build_integral(image, B, H, W, d, st));
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
    return dense_bad_at_kpts_launch(image, B, H, W, kpts, K, pair_table, P, desc_mode, temperature, normalize, desc, ws,
                                    ws_bytes, (cudaStream_t)stream);
}

extern "C" int om_gather_descriptors_f32(const float* desc_map, int B, int D, int H, int W, const float* kpts, int K,
                                         int subpixel, float* desc, void* stream) {
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
    if (image == nullptr || kpts == nullptr || desc == nullptr) return OM_ERR_NULL;
    if (B <= 0 || H <= 1 || W <= 1 || K <= 0) return OM_ERR_SHAPE;
    OM_TRY(check_table(pair_table, P));
    if (ws == nullptr || ws_bytes < dense_bad_workspace_bytes(B, H, W)) return OM_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const DenseWs d = carve_dense(ws, B, H, W);
    if (stage == 0) return build_integral(image, B, H, W, d, st);
    DenseKpArgs a{};
    a.I = d.I; a.B = B; a.H = H; a.W = W; a.kpts = kpts; a.K = K; a.table = pair_table; a.P = P; a.mode = desc_mode;
    a.temperature = temperature; a.normalize = normalize; a.desc = desc;
    constexpr int GROUPS = 4;
    const long long total = (long long)B * K;
    dense_at_kpts_kernel<GROUPS><<<(unsigned)((total + GROUPS - 1) / GROUPS), GROUPS * TPG,
                                   (size_t)GROUPS * 47 * 47 * sizeof(float), st>>>(a);
    OM_AFTER_LAUNCH();
    return OM_OK;
}
