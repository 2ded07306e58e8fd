// stand-alone probe: TMA 3D box load with the tensor map as a __grid_constant__ parameter vs in global memory
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void probe(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, unsigned* out, int x0, int y0, int z, int bw) {
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ __align__(8) unsigned long long bar;
    unsigned* win = reinterpret_cast<unsigned*>(raw);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bw * 48 * 4) : "memory");
        const CUtensorMap* m = MODE == 0 ? &pmap : gmap;
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                smem_u32(win)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(&bar)), "r"(x0), "r"(y0), "r"(z) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(
            smem_u32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 48 * 48; i += blockDim.x) out[i] = win[(i / 48) * bw + (i % 48 < bw ? i % 48 : 0)];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1; const unsigned BW = argc > 2 ? atoi(argv[2]) : 48;
    const int IP = 208, Hi = 167, B = 2;
    std::vector<unsigned> h((size_t)IP * Hi * B);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned)i;
    unsigned *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&o, 48 * 48 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry point: err %d q %d fn %p\n", (int)e, (int)q, fn);
    CUtensorMap map;
    const cuuint64_t dims[3] = {IP, Hi, B};
    const cuuint64_t strides[2] = {(cuuint64_t)IP * 4, (cuuint64_t)IP * 4 * Hi};
    const cuuint32_t box[3] = {BW, 48, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    CUtensorMap* gm;
    cudaMalloc(&gm, sizeof(map));
    cudaMemcpy(gm, &map, sizeof(map), cudaMemcpyHostToDevice);
    std::vector<unsigned> ho(48 * 48);
    for (int mode = 0; mode < 2; ++mode) {
        if (only >= 0 && mode != only) continue;
        cudaMemset(o, 0, 48 * 48 * 4);
        if (mode == 0) probe<0><<<1, 64, 64 * 48 * 4>>>(map, gm, o, 5, 7, 1, (int)BW);
        else probe<1><<<1, 64, 64 * 48 * 4>>>(map, gm, o, 5, 7, 1, (int)BW);
        e = cudaDeviceSynchronize();
        printf("mode %d: sync err %d (%s)\n", mode, (int)e, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        cudaMemcpy(ho.data(), o, 48 * 48 * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int y = 0; y < 48; ++y)
            for (int x = 0; x < 48; ++x) bad += x < (int)BW && ho[y * 48 + x] != (unsigned)((size_t)1 * IP * Hi + (size_t)(7 + y) * IP + 5 + x);
        printf("mode %d: mismatches %d, first %u\n", mode, bad, ho[0]);
    }
    return 0;
}
