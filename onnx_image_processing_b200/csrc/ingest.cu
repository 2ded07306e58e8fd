// Camera-frame ingest: BGR (or grey) uint8 frame -> grey -> bilinear resize -> uint8 / float32 model input, one kernel.
// Replaces the host preprocessing of the reference's real callers, sample/visual_odometry.py:65-92 (load_image_from_array:
// cv2.cvtColor(BGR2GRAY) + cv2.resize(INTER_LINEAR) + astype(float32)), with OpenCV's own integer arithmetic so that the
// pixels the matcher sees are the ones the reference's pipeline would have produced:
//   grey   = (B * 3735 + G * 19235 + R * 9798 + 2^14) >> 15                       (OpenCV's 15-bit fixed-point luma)
//   resize : source position fx = (dx + 0.5) * scale - 0.5 evaluated in double and rounded to float, sx = floor(fx), weights
//            rounded to 11 bits (cvRound((1 - fx) * 2048), cvRound(fx * 2048)); horizontally an out-of-range position moves to
//            the border pixel with weight 0, vertically the fractional weight is kept and only the two row indices are clipped;
//            horizontal pass in 32-bit integers, vertical pass (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2
// (identical to cv2 4.13 for every size tried, shrinking, enlarging and mixed: tests/golden/make_golden_f4.py).
// The uint8 result feeds om_match_pairs(OM_IMAGE_U8) directly: the frame is read once, 1 byte per model pixel is written.
#include "common.cuh"

namespace om {

namespace {

struct IngestArgs {
    const unsigned char* src;
    int Hin, Win, C, Hout, Wout;
    double scale_x, scale_y;
    unsigned char* d8;
    float* df;
};

// source index and the two 11-bit weights of output coordinate d (cv::resize, INTER_LINEAR, 8-bit)
template <bool CLAMP_WEIGHT>
__device__ __forceinline__ void linear_coef(int d, double scale, int n_in, int& s, int& w0, int& w1) {
    const double t = __dadd_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), -0.5);   // no contraction: same roundings as the host code
    float f = __double2float_rn(t);
    int si = (int)floorf(f);
    f = __fsub_rn(f, (float)si);
    if (CLAMP_WEIGHT) {
        if (si < 0) { f = 0.0f; si = 0; }
        if (si >= n_in - 1) { f = 0.0f; si = n_in - 1; }
    }
    w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));
    w1 = __float2int_rn(__fmul_rn(f, 2048.0f));
    s = si;
}

__device__ __forceinline__ int grey_at(const unsigned char* row, int x, int C) {
    if (C == 1) return (int)__ldg(row + x);
    const unsigned char* p = row + 3 * (size_t)x;
    const int b = __ldg(p), g = __ldg(p + 1), r = __ldg(p + 2);
    return (b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15;
}

__global__ void __launch_bounds__(256) ingest_kernel(IngestArgs a) {
    const int x0 = 4 * (blockIdx.x * 64 + threadIdx.x);
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int z = blockIdx.z;
    if (x0 >= a.Wout || y >= a.Hout) return;
    int sy, b0, b1;
    linear_coef<false>(y, a.scale_y, a.Hin, sy, b0, b1);
    const size_t pitch = (size_t)a.Win * a.C;
    const unsigned char* r0 = a.src + ((size_t)z * a.Hin + clampi(sy, 0, a.Hin - 1)) * pitch;
    const unsigned char* r1 = a.src + ((size_t)z * a.Hin + clampi(sy + 1, 0, a.Hin - 1)) * pitch;
    int out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = min(x0 + j, a.Wout - 1);
        int sx, a0, a1;
        linear_coef<true>(x, a.scale_x, a.Win, sx, a0, a1);
        const int sx1 = min(sx + 1, a.Win - 1);
        const int S0 = grey_at(r0, sx, a.C) * a0 + grey_at(r0, sx1, a.C) * a1;
        const int S1 = grey_at(r1, sx, a.C) * a0 + grey_at(r1, sx1, a.C) * a1;
        const int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
        out[j] = min(max(v, 0), 255);
    }
    const size_t o = ((size_t)z * a.Hout + y) * a.Wout + x0;
    const bool full = x0 + 3 < a.Wout;
    if (a.d8 != nullptr) {
        if (full && (a.Wout & 3) == 0 && (reinterpret_cast<uintptr_t>(a.d8) & 3) == 0) {
            *reinterpret_cast<unsigned int*>(a.d8 + o) =
                (unsigned)out[0] | ((unsigned)out[1] << 8) | ((unsigned)out[2] << 16) | ((unsigned)out[3] << 24);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (x0 + j < a.Wout) a.d8[o + j] = (unsigned char)out[j];
        }
    }
    if (a.df != nullptr) {
        if (full && (a.Wout & 3) == 0 && (reinterpret_cast<uintptr_t>(a.df) & 15) == 0) {
            *reinterpret_cast<float4*>(a.df + o) = make_float4((float)out[0], (float)out[1], (float)out[2], (float)out[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (x0 + j < a.Wout) a.df[o + j] = (float)out[j];
        }
    }
}

}  // namespace

}  // namespace om

using namespace om;

extern "C" int om_preprocess_u8(const unsigned char* src, int B, int Hin, int Win, int channels, int Hout, int Wout,
                                unsigned char* dst_u8, float* dst_f32, void* stream) {
    OM_ON_DEVICE_OF(src);
    if (src == nullptr || (dst_u8 == nullptr && dst_f32 == nullptr)) return OM_ERR_NULL;
    if (B <= 0 || Hin <= 0 || Win <= 0 || Hout <= 0 || Wout <= 0) return OM_ERR_SHAPE;
    if (channels != 1 && channels != 3) return OM_ERR_PARAM;
    if (B > 65535 || (long long)Hin * Win >= (1ll << 29) || (long long)Hout * Wout >= (1ll << 31)) return OM_ERR_LIMIT;
    IngestArgs a{};
    a.src = src; a.Hin = Hin; a.Win = Win; a.C = channels; a.Hout = Hout; a.Wout = Wout;
    // cv::resize: inv_scale = dsize / ssize in double, scale = 1 / inv_scale
    a.scale_x = 1.0 / ((double)Wout / (double)Win);
    a.scale_y = 1.0 / ((double)Hout / (double)Hin);
    a.d8 = dst_u8; a.df = dst_f32;
    const dim3 block(64, 4), grid((unsigned)((Wout + 255) / 256), (unsigned)((Hout + 3) / 4), (unsigned)B);
    ingest_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(a);
    OM_AFTER_LAUNCH();
    return OM_OK;
}
