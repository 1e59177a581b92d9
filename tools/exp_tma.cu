// Development micro-benchmark: how fast can persistent CTAs stream a row-major fp32 matrix
// through TMA tile loads (no compute)?  Variants: box rows, boxes per stage (one mbarrier per
// stage), ring depth, swizzle mode.  Built into tools/exp_tma (standalone binary).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}

struct P { int64_t n_blocks; int block_rows, box_rows, box_cols, f, boxes_per_stage, stages; };

// A "block" = block_rows rows x f cols; it is loaded as (block_rows/box_rows) x (f/box_cols) boxes,
// column-major order within a row group; boxes_per_stage consecutive boxes share a barrier.
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap map, const P p) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const uint32_t box_bytes = p.box_rows * p.box_cols * 4, stage_bytes = box_bytes * p.boxes_per_stage;
    const uint32_t ring = smem_u32(smem), bars = ring + p.stages * stage_bytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(bars + 16 * s, 1); mbar_init(bars + 16 * s + 8, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t nb = (p.n_blocks - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const int rgroups = p.block_rows / p.box_rows, cboxes = p.f / p.box_cols;
    const int boxes = rgroups * cboxes, stages_per_block = boxes / p.boxes_per_stage;
    if (threadIdx.x == 0) {
        uint32_t it = 0;
        for (int64_t i = 0; i < nb; ++i) {
            const int row0 = (int)((blockIdx.x + i * gridDim.x) * p.block_rows);
            for (int s = 0; s < stages_per_block; ++s, ++it) {
                const int st = it % p.stages;
                mbar_wait(bars + 16 * st + 8, ((it / p.stages) & 1) ^ 1);
                mbar_expect_tx(bars + 16 * st, stage_bytes);
                for (int b = 0; b < p.boxes_per_stage; ++b) {
                    const int box = s * p.boxes_per_stage + b;
                    const int rg = box / cboxes, c = box % cboxes;
                    tma_load_2d(ring + st * stage_bytes + b * box_bytes, &map, c * p.box_cols, row0 + rg * p.box_rows, bars + 16 * st);
                }
            }
        }
    } else if (threadIdx.x == 32) {
        uint32_t it = 0;
        for (int64_t i = 0; i < nb * stages_per_block; ++i, ++it) {
            const int st = it % p.stages;
            mbar_wait(bars + 16 * st, (it / p.stages) & 1);
            mbar_arrive(bars + 16 * st + 8);
        }
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 4000000;
    const int f = 512;
    float* X;
    cudaMalloc(&X, n * f * 4);
    cudaMemset(X, 0, n * f * 4);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)fp;
    struct V { int block_rows, box_rows, box_cols, bps, stages; CUtensorMapSwizzle sw; CUtensorMapDataType dt; const char* name; };
    std::vector<V> vs = {
        {64, 64, 32, 1, 11, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, "64x32 box, 1/stage, 11 st (current P1)"},
        {64, 64, 32, 1, 24, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, "64x32 box, 1/stage, 24 st"},
        {64, 64, 32, 4, 6, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, "64x32 box, 4/stage, 6 st (192KB)"},
        {64, 64, 32, 4, 3, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, "64x32 box, 4/stage, 3 st (96KB)"},
        {64, 64, 32, 16, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, "64x32 box, 16/stage, 1 st"},
        {128, 128, 32, 1, 12, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, "128x32 box, 12 st (192KB)"},
        {256, 256, 32, 1, 6, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, "256x32 box, 6 st (192KB)"},
        {64, 16, 32, 4, 24, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, "16x32 box, 4/stage"},
        {64, 64, 32, 1, 24, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, "64x32 box f32 type, 24 st"},
        {64, 64, 32, 1, 24, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, "64x32 box SW128_32B, 24 st"},
        {64, 64, 32, 1, 24, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, "64x32 box no swizzle, 24 st"},
        {64, 8, 256, 1, 24, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, "8x256 box (1KB rows) no swizzle, 24 st"},
        {64, 32, 256, 1, 6, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, "32x256 box (1KB rows) no swizzle, 6 st"},
    };
    for (auto& v : vs) {
        CUtensorMap map;
        cuuint64_t dims[2] = {(cuuint64_t)f, (cuuint64_t)n}, strides[1] = {(cuuint64_t)f * 4};
        cuuint32_t box[2] = {(cuuint32_t)v.box_cols, (cuuint32_t)v.box_rows}, el[2] = {1, 1};
        CUresult rc = enc(&map, v.dt, 2, X, dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE, v.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc) { printf("%s: encode failed %d\n", v.name, rc); continue; }
        P p{n / v.block_rows, v.block_rows, v.box_rows, v.box_cols, f, v.bps, v.stages};
        size_t smem = (size_t)v.stages * v.bps * v.box_rows * v.box_cols * 4 + v.stages * 16 + 1024;
        if (cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { printf("%s: smem %zu too large\n", v.name, smem); cudaGetLastError(); continue; }
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        stream_kernel<<<148, 64, smem>>>(map, p);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int r = 0; r < 3; ++r) stream_kernel<<<148, 64, smem>>>(map, p);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
        cudaError_t err = cudaGetLastError();
        printf("%-45s smem %3zu KB  %7.3f ms  %7.1f GB/s %s\n", v.name, smem / 1024, ms, (double)n * f * 4 / ms / 1e6, err ? cudaGetErrorString(err) : "");
    }
    return 0;
}
