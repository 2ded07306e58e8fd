// Issue-rate probe for sm_100a: scalar FFMA / FADD / FMUL against the packed f32x2 forms (FFMA2 / FADD2 / FMUL2), with
// independent accumulator chains so that the pipe, not the dependency latency, is measured.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_probe f32x2_probe.cu && ./f32x2_probe
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int CH = 16;   // independent chains per thread (floats); packed form uses CH/2 float2 chains

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(float* out, float s, long long* cyc) {
    float a[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) a[i] = (float)(threadIdx.x + i) * 1e-3f;
    const float2 s2 = make_float2(s, s * 0.999f);
    const float2 t2 = make_float2(1e-7f, 2e-7f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < CH; ++i) a[i] = fmaf(a[i], s2.x, t2.x);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < CH; i += 2) {
                float2 v = make_float2(a[i], a[i + 1]);
                v = __ffma2_rn(v, s2, t2);
                a[i] = v.x; a[i + 1] = v.y;
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < CH; ++i) a[i] = __fadd_rn(a[i], t2.x);
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < CH; i += 2) {
                float2 v = make_float2(a[i], a[i + 1]);
                v = __fadd2_rn(v, t2);
                a[i] = v.x; a[i + 1] = v.y;
            }
        } else if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < CH; ++i) a[i] = __fmul_rn(a[i], s2.x);
        } else if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < CH; i += 2) {
                float2 v = make_float2(a[i], a[i + 1]);
                v = __fmul2_rn(v, s2);
                a[i] = v.x; a[i + 1] = v.y;
            }
        } else if (MODE == 6) {   // FFMA with three distinct register operands (no immediate / uniform)
#pragma unroll
            for (int i = 0; i < CH; ++i) a[i] = fmaf(a[i], a[(i + 1) % CH], a[(i + 2) % CH]);
        } else if (MODE == 7) {
#pragma unroll
            for (int i = 0; i < CH; i += 2) {
                float2 v = make_float2(a[i], a[i + 1]);
                float2 w = make_float2(a[(i + 2) % CH], a[(i + 3) % CH]);
                float2 u = make_float2(a[(i + 4) % CH], a[(i + 5) % CH]);
                v = __ffma2_rn(v, w, u);
                a[i] = v.x; a[i + 1] = v.y;
            }
        }
    }
    const long long t1 = clock64();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// shared-memory read bandwidth: every thread LDS.128 from conflict-free addresses
__global__ void __launch_bounds__(512, 1) lds_probe(float* out, long long* cyc, int iters) {
    extern __shared__ float4 sm4[];
    for (int i = threadIdx.x; i < 8192; i += 512) sm4[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 v = sm4[k * 512 + threadIdx.x];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, float flop_per_op) {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4);
    cudaMalloc(&cyc, 148 * 8);
    probe<MODE><<<148, 512>>>(out, 1.0001f, cyc);
    probe<MODE><<<148, 512>>>(out, 1.0001f, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < 148; ++i) c += h[i];
    c /= 148;
    const double lane_ops = (double)ITERS * CH * 512;      // float results per CTA
    printf("%-28s %9.0f cyc  -> %6.1f float results / clk / SM (%s)\n", name, c, lane_ops / c, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("FFMA  (reg, uniform, uniform)", 2);
    run<1>("FFMA2 (reg, uniform, uniform)", 2);
    run<2>("FADD", 1);
    run<3>("FADD2", 1);
    run<4>("FMUL", 1);
    run<5>("FMUL2", 1);
    run<6>("FFMA  (3 registers)", 2);
    run<7>("FFMA2 (3 register pairs)", 2);
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    cudaFuncSetAttribute(lds_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    lds_probe<<<148, 512, 131072>>>(out, cyc, 256);
    lds_probe<<<148, 512, 131072>>>(out, cyc, 256);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < 148; ++i) c += h[i];
    c /= 148;
    printf("LDS.128 conflict-free: %.0f cyc for %d bytes -> %.1f B/clk/SM (%s)\n", c, 256 * 16 * 512 * 16, 256.0 * 16 * 512 * 16 / c,
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}
