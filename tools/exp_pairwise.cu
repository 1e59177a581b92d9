// Round-2 experiment (not built by build(), not part of the library): register tiling of the
// pairwise Chebyshev-gap kernel (csrc/prune.cu).  Round 1 measured 27.7 ms for 128 binned columns
// x 10 M rows with 4 x 4 tiles (2 LDS.128 per 16 pairs: shared-memory bound; ALU bound ~5 ms).
// Variants, each checked against a naive kernel on the same random bins:
//   0  4 x 4 fp32 tiles (the committed kernel's inner loop)
//   1  8 x 8 fp32 tiles (4 LDS.128 per 64 pairs)
//   2  8 x 8 tiles on half2 = two ROWS per register (bins < 2048 are exact in fp16):
//      HADD2 + HMNMX2 process two rows per instruction and the chunk takes half the shared memory
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp_pairwise tools/exp_pairwise.cu
// Run:   ./tools/exp_pairwise [rows, default 10000000] [columns, default 128]
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define RT(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            printf("FAIL %s -> %s at line %d\n", #x, cudaGetErrorString(e_), __LINE__); \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

__global__ void naive_kernel(const int* bins, long n, int d, int* gap) {
    const int i = blockIdx.x, j = blockIdx.y;
    if (j <= i) return;
    int m = 0;
    for (long r = threadIdx.x; r < n; r += blockDim.x)
        m = max(m, abs(bins[(long)i * n + r] - bins[(long)j * n + r]));
    atomicMax(&gap[i * d + j], m);
    atomicMax(&gap[j * d + i], m);
}

__global__ void fill_kernel(int* bins, long total, unsigned seed) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        unsigned x = (unsigned)i * 2654435761u ^ seed;
        x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
        bins[i] = (int)(x % 24u);
    }
}

// ---- fp32 tiles: TS x TS pairs per thread, chunk [R][stride] floats --------------------------------
template <int TS>
__global__ void __launch_bounds__(256, 1)
tile_f32_kernel(const int* __restrict__ bins, long n, int d, int T, int R, int n_pair_tiles,
                int* __restrict__ gap) {
    extern __shared__ float chunk[];
    const int stride = TS * T + 4;
    // pair tile of this thread: triangular index -> (ti, tj), ti <= tj
    const int pt = blockIdx.y * 256 + threadIdx.x;
    int ti = 0, rem = pt;
    while (ti < T && rem >= T - ti) { rem -= T - ti; ++ti; }
    const int tj = ti + rem;
    const bool active = pt < n_pair_tiles;
    float m[TS][TS];
#pragma unroll
    for (int x = 0; x < TS; ++x)
#pragma unroll
        for (int y = 0; y < TS; ++y) m[x][y] = 0.f;
    const long n_chunks = (n + R - 1) / R;
    for (long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        const long r0 = ch * R;
        __syncthreads();
        for (int idx = threadIdx.x; idx < TS * T * R; idx += 256) {
            const int c = idx / R, r = idx - c * R;
            chunk[r * stride + c] = (c < d && r0 + r < n) ? (float)bins[(long)c * n + r0 + r] : 0.f;
        }
        __syncthreads();
        if (!active) continue;
        const float* pa = chunk + TS * ti;
        const float* pb = chunk + TS * tj;
#pragma unroll 2
        for (int r = 0; r < R; ++r) {
            float a[TS], b[TS];
#pragma unroll
            for (int v = 0; v < TS / 4; ++v) {
                const float4 fa = *reinterpret_cast<const float4*>(pa + r * stride + 4 * v);
                const float4 fb = *reinterpret_cast<const float4*>(pb + r * stride + 4 * v);
                a[4 * v] = fa.x; a[4 * v + 1] = fa.y; a[4 * v + 2] = fa.z; a[4 * v + 3] = fa.w;
                b[4 * v] = fb.x; b[4 * v + 1] = fb.y; b[4 * v + 2] = fb.z; b[4 * v + 3] = fb.w;
            }
#pragma unroll
            for (int x = 0; x < TS; ++x)
#pragma unroll
                for (int y = 0; y < TS; ++y) m[x][y] = fmaxf(m[x][y], fabsf(a[x] - b[y]));
        }
    }
    if (!active) return;
#pragma unroll
    for (int x = 0; x < TS; ++x)
#pragma unroll
        for (int y = 0; y < TS; ++y) {
            const int i = TS * ti + x, j = TS * tj + y;
            if (i < d && j < d && i != j && m[x][y] > 0.f) {
                atomicMax(&gap[i * d + j], (int)m[x][y]);
                atomicMax(&gap[j * d + i], (int)m[x][y]);
            }
        }
}

// ---- half2 tiles: 8 x 8 pairs, two rows per register, chunk [R / 2][stride] half2 ---------------------
__global__ void __launch_bounds__(256, 1)
tile_h2_kernel(const int* __restrict__ bins, long n, int d, int T, int R, int n_pair_tiles,
               int* __restrict__ gap) {
    extern __shared__ __half2 chunk_h[];
    constexpr int TS = 8;
    const int stride = TS * T + 4;                 // half2 elements per packed row pair
    const int pt = blockIdx.y * 256 + threadIdx.x;
    int ti = 0, rem = pt;
    while (ti < T && rem >= T - ti) { rem -= T - ti; ++ti; }
    const int tj = ti + rem;
    const bool active = pt < n_pair_tiles;
    __half2 m[TS][TS];
#pragma unroll
    for (int x = 0; x < TS; ++x)
#pragma unroll
        for (int y = 0; y < TS; ++y) m[x][y] = __float2half2_rn(0.f);
    const long n_chunks = (n + R - 1) / R;
    for (long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        const long r0 = ch * R;
        __syncthreads();
        for (int idx = threadIdx.x; idx < TS * T * (R / 2); idx += 256) {
            const int c = idx / (R / 2), rp = idx - c * (R / 2);
            const long ra = r0 + 2 * rp;
            const float lo = (c < d && ra < n) ? (float)bins[(long)c * n + ra] : 0.f;
            const float hi = (c < d && ra + 1 < n) ? (float)bins[(long)c * n + ra + 1] : 0.f;
            chunk_h[rp * stride + c] = __floats2half2_rn(lo, hi);
        }
        __syncthreads();
        if (!active) continue;
        const __half2* pa = chunk_h + TS * ti;
        const __half2* pb = chunk_h + TS * tj;
#pragma unroll 2
        for (int rp = 0; rp < R / 2; ++rp) {
            __half2 a[TS], b[TS];
#pragma unroll
            for (int v = 0; v < TS / 4; ++v) {
                const uint4 ua = *reinterpret_cast<const uint4*>(pa + rp * stride + 4 * v);
                const uint4 ub = *reinterpret_cast<const uint4*>(pb + rp * stride + 4 * v);
                a[4 * v] = *reinterpret_cast<const __half2*>(&ua.x);
                a[4 * v + 1] = *reinterpret_cast<const __half2*>(&ua.y);
                a[4 * v + 2] = *reinterpret_cast<const __half2*>(&ua.z);
                a[4 * v + 3] = *reinterpret_cast<const __half2*>(&ua.w);
                b[4 * v] = *reinterpret_cast<const __half2*>(&ub.x);
                b[4 * v + 1] = *reinterpret_cast<const __half2*>(&ub.y);
                b[4 * v + 2] = *reinterpret_cast<const __half2*>(&ub.z);
                b[4 * v + 3] = *reinterpret_cast<const __half2*>(&ub.w);
            }
#pragma unroll
            for (int x = 0; x < TS; ++x)
#pragma unroll
                for (int y = 0; y < TS; ++y) m[x][y] = __hmax2(m[x][y], __habs2(__hsub2(a[x], b[y])));
        }
    }
    if (!active) return;
#pragma unroll
    for (int x = 0; x < TS; ++x)
#pragma unroll
        for (int y = 0; y < TS; ++y) {
            const int i = TS * ti + x, j = TS * tj + y;
            const int v = (int)fmaxf(__low2float(m[x][y]), __high2float(m[x][y]));
            if (i < d && j < d && i != j && v > 0) {
                atomicMax(&gap[i * d + j], v);
                atomicMax(&gap[j * d + i], v);
            }
        }
}

int main(int argc, char** argv) {
    const long n = argc > 1 ? atol(argv[1]) : 10000000;
    const int d = argc > 2 ? atoi(argv[2]) : 128;
    int* bins = nullptr;
    int *gap_ref = nullptr, *gap = nullptr;
    RT(cudaMalloc(&bins, (size_t)n * d * sizeof(int)));
    RT(cudaMalloc(&gap_ref, (size_t)d * d * sizeof(int)));
    RT(cudaMalloc(&gap, (size_t)d * d * sizeof(int)));
    fill_kernel<<<148 * 8, 256>>>(bins, n * d, 12345u);
    RT(cudaMemset(gap_ref, 0, (size_t)d * d * sizeof(int)));
    naive_kernel<<<dim3(d, d), 256>>>(bins, n, d, gap_ref);
    RT(cudaDeviceSynchronize());
    std::vector<int> h_ref((size_t)d * d), h((size_t)d * d);
    RT(cudaMemcpy(h_ref.data(), gap_ref, h_ref.size() * sizeof(int), cudaMemcpyDeviceToHost));
    int max_smem = 0;
    RT(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0));
    cudaEvent_t e0, e1;
    RT(cudaEventCreate(&e0));
    RT(cudaEventCreate(&e1));
    for (int variant = 0; variant < 3; ++variant) {
        const int TS = variant == 0 ? 4 : 8;
        const int T = (d + TS - 1) / TS;
        const int n_pair_tiles = T * (T + 1) / 2;
        const int stride = TS * T + 4;
        const int elem = variant == 2 ? 4 /* half2 = two rows */ : 4;
        int R = (max_smem - 1024) / (stride * elem);          // packed rows (half2) or rows (fp32)
        if (variant == 2) R *= 2;
        R = R > 512 ? 512 : R;
        R -= R % 16;
        const size_t smem = variant == 2 ? (size_t)(R / 2) * stride * 4 : (size_t)R * stride * 4;
        const int batches = (n_pair_tiles + 255) / 256;
        const dim3 grid(148 / batches > 0 ? 148 / batches : 1, batches);
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            RT(cudaMemset(gap, 0, (size_t)d * d * sizeof(int)));
            RT(cudaEventRecord(e0));
            if (variant == 0) {
                RT(cudaFuncSetAttribute(tile_f32_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                tile_f32_kernel<4><<<grid, 256, smem>>>(bins, n, d, T, R, n_pair_tiles, gap);
            } else if (variant == 1) {
                RT(cudaFuncSetAttribute(tile_f32_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                tile_f32_kernel<8><<<grid, 256, smem>>>(bins, n, d, T, R, n_pair_tiles, gap);
            } else {
                RT(cudaFuncSetAttribute(tile_h2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                tile_h2_kernel<<<grid, 256, smem>>>(bins, n, d, T, R, n_pair_tiles, gap);
            }
            RT(cudaEventRecord(e1));
            RT(cudaEventSynchronize(e1));
            RT(cudaGetLastError());
            RT(cudaEventElapsedTime(&ms, e0, e1));
        }
        RT(cudaMemcpy(h.data(), gap, h.size() * sizeof(int), cudaMemcpyDeviceToHost));
        long bad = 0;
        for (size_t k = 0; k < h.size(); ++k) bad += h[k] != h_ref[k];
        printf("variant %d (%s): R = %d, smem = %zu, grid = %u x %u, %.3f ms, %.3e pair-rows/s, "
               "mismatches %ld\n", variant,
               variant == 0 ? "4x4 fp32" : variant == 1 ? "8x8 fp32" : "8x8 half2 (2 rows/reg)",
               R, smem, grid.x, grid.y, ms, (double)d * (d - 1) / 2 * n / ms * 1e3, bad);
    }
    return 0;
}
