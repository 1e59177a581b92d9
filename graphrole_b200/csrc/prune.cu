// Feature pruning between ReFeX recursion levels on the device ("next" #1).
//
// The reference re-codes every feature column into vertical logarithmic bins
// (graphrole/features/prune.py:13-56: np.unique + one masking pass per bin) and links two columns
// when their binned versions differ by at most a threshold everywhere (Chebyshev pdist,
// prune.py:104-108).  Both steps are O(n) per column / column pair and run between every pair of
// recursion levels (extract.py:135-137).  Here:
//
//   binning   per column: cub radix sort of the VALUES alone; one warp walks the bin boundaries
//             (each boundary = end of the run of equal values that contains sorted position
//             want-1, found by a 32-ary search) -- the same integers the reference gets from
//             searchsorted(cumsum(counts), binned_len + bin_size) -- and records every bin's
//             largest value.  Bins end on run ends, so a value's bin is the number of bins whose
//             largest value lies below it: every row looks its own value up among those (<= a
//             few dozen) thresholds, reading and writing in row order.  (Round 1 sorted
//             (value, row) pairs and scattered the bin of every sorted position back to its
//             row: twice the sort traffic plus n random 4-byte stores per column.)
//             Columns are pulled out of the row-major feature matrix 32 (fp32) / 16 (fp64) at a
//             time through a shared-memory transpose, so the matrix is read once, coalesced.
//   distance  persistent CTAs stream row chunks of ALL binned columns through shared memory and
//             every thread keeps the running max |a - b| of a 4 x 4 block of column pairs in
//             registers (bins are small integers, exact in fp32: |a - b| folds into FADD's
//             operand modifiers + FMNMX); one atomicMax per pair per CTA at the end.
//
// Bound: binning is sort-bound (cub, ~8 bytes moved per key per pass); the distance kernel is
// ALU-bound at 2 instructions per (pair, row) -- F^2/2 * n * 2 lane-ops, e.g. F = 128, n = 10 M:
// 1.6e11 lane-ops ~ 5 ms on 148 SMs -- while reading the binned matrix (F * n * 4 bytes) once per
// batch of 256 pair tiles.

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <new>

#include "common.cuh"

using namespace gr;

struct gr_pruner {
    int device = 0;
    int64_t n = 0;
    int sms = 148;
    int max_smem = 0;
    void* colbuf = nullptr;      // [32][n] fp32 or [16][n] fp64: the column group being binned
    void* keys_sorted = nullptr; // [n] of the key type (8 bytes per entry reserved)
    void* tops = nullptr;        // [n + 1] of the key type: largest value of every bin
    int32_t* bounds = nullptr;   // [n + 1] bin boundaries (exclusive end positions), ascending
    int32_t* n_bounds = nullptr; // [1]
    void* cub_temp = nullptr;
    size_t cub_temp_bytes = 0;
    int2* pair_tiles = nullptr;  // (ti, tj) list of the distance kernel, ti <= tj
    int pair_tiles_for_d = -1;
};

namespace {

constexpr int kGroupBytes = 128;    // columns pulled per pass = 128 / sizeof(key)

// ---- column extraction: row-major [n, d] -> column-major group [G][n] -----------------------------
template <typename T>
__global__ void extract_columns_kernel(const T* __restrict__ X, int64_t ldx, int64_t n, int c0,
                                       int cols, T* __restrict__ out) {
    constexpr int G = kGroupBytes / (int)sizeof(T);
    __shared__ T tile[32][G + 1];
    const int tx = threadIdx.x % G, ty = threadIdx.x / G;       // 256 threads: 256 / G rows per pass
    constexpr int RPP = 256 / G;
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    for (int rr = ty; rr < 32; rr += RPP) {
        const int64_t r = r0 + rr;
        tile[rr][tx] = (r < n && tx < cols) ? X[r * ldx + c0 + tx] : T(0);
    }
    __syncthreads();
    // write: lanes along rows
    const int lr = threadIdx.x % 32, lc0 = threadIdx.x / 32;    // 8 columns per pass
    for (int c = lc0; c < cols; c += 8) {
        const int64_t r = r0 + lr;
        if (r < n) out[(int64_t)c * n + r] = tile[lr][c];
    }
}

// ---- bin boundaries: one warp, 32-ary upper-bound searches ----------------------------------------
// Follows prune.py:36-54: bin_size = max(int(frac * unbinned), 1); the bin ends with the run of
// equal values containing sorted position binned + bin_size - 1.
template <typename T>
__global__ void bin_bounds_kernel(const T* __restrict__ keys, int64_t n, double frac,
                                  int32_t* __restrict__ bounds, T* __restrict__ tops,
                                  int32_t* __restrict__ n_bounds) {
    const int lane = threadIdx.x;
    int64_t done = 0;
    int nb = 0;
    while (done < n) {
        const int64_t size = max((int64_t)(frac * (double)(n - done)), (int64_t)1);
        const int64_t want = done + size;                 // >= 1, <= n
        const T v = keys[want - 1];
        // first position p in [want, n] with p == n or keys[p] > v (keys ascending; -0.0 == 0.0)
        int64_t lo = want, hi = n;                        // answer in [lo, hi]
        {
            // galloping first round: probes at want + 2^lane - 1.  A column of distinct values
            // ends the search here (lane 0 already sees a larger key) -- one dependent load per
            // bin instead of log32(n) rounds; a tie of length t leaves a range of <= t to search.
            const int64_t off = lane < 31 ? (((int64_t)1 << lane) - 1) : (int64_t)1 << 40;
            const int64_t p = want + off;
            const bool greater = p < n ? (keys[p] > v) : true;
            const unsigned m = __ballot_sync(0xffffffffu, greater);     // lane 31 always votes
            const int first = __ffs(m) - 1;
            const int64_t pf = want + (first < 31 ? (((int64_t)1 << first) - 1) : (int64_t)1 << 40);
            if (first > 0) lo = want + (((int64_t)1 << (first - 1)) - 1) + 1;
            hi = pf < n ? pf : n;
        }
        while (lo < hi) {
            const int64_t span = hi - lo;
            const int64_t step = (span + 31) / 32;        // 32 ascending probes cover [lo, hi)
            const int64_t p = lo + (int64_t)(lane + 1) * step - 1;
            const bool greater = p < hi ? (keys[p] > v) : true;
            const unsigned m = __ballot_sync(0xffffffffu, greater);
            if (m == 0) {                                 // every probe <= v: answer is past them
                lo = min(lo + 32 * step, hi);
                continue;
            }
            const int first = __ffs(m) - 1;               // answer in (probe[first-1], probe[first]]
            const int64_t pf = lo + (int64_t)(first + 1) * step - 1;
            lo = lo + (int64_t)first * step;
            hi = pf < hi ? pf : hi;
        }
        done = lo;
        if (lane == 0) {
            bounds[nb] = (int32_t)done;
            tops[nb] = v;          // the bin holds exactly the values in (tops[nb - 1], v]
        }
        ++nb;
    }
    if (lane == 0) *n_bounds = nb;
}

// ---- bin of every row: number of bins whose largest value is below the row's value --------------
template <typename T>
__global__ void assign_bins_kernel(const T* __restrict__ values, const T* __restrict__ tops,
                                   const int32_t* __restrict__ n_bounds, int64_t n,
                                   int32_t* __restrict__ bins_col) {
    __shared__ T st[512];
    const int nb = *n_bounds;
    const int cached = min(nb, 512);
    for (int i = threadIdx.x; i < cached; i += blockDim.x) st[i] = tops[i];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const T x = values[r];
    int lo = 0, hi = nb;           // first bin whose largest value is >= x (== compares -0.0, 0.0)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const T t = mid < cached ? st[mid] : tops[mid];
        if (t < x) lo = mid + 1; else hi = mid;
    }
    bins_col[r] = lo;
}

// ---- pairwise Chebyshev distance of binned columns --------------------------------------------------
__global__ void pair_tile_table_kernel(int2* table, int T) {
    const int ti = blockIdx.x, tj = threadIdx.x + blockIdx.y * blockDim.x;
    if (ti < T && tj < T && ti <= tj)
        table[ti * T - ti * (ti - 1) / 2 + (tj - ti)] = make_int2(ti, tj);
}

// bins: column-major [d][ldb] int32.  smem chunk: [R][stride] fp32, stride = 4 * T4 columns
// padded to an odd number of float4 (float4 alignment, conflict-free transposing stores).
// grid.x = persistent chunk workers, grid.y = batches of pair tiles.  A CTA's 256 threads are
// P pair slots x S row slices (P = power of two >= min(#pair tiles, 256)): pair tile
// pt = blockIdx.y + slot * gridDim.y, rows slice, slice + S, ... of every chunk.  A thread keeps
// the running maxima of a TS x TS block of column pairs in registers: TS = 4 (2 LDS.128 per 16
// pairs) for narrow frames, TS = 8 (4 LDS.128 per 64 pairs: measured 29.9 -> 18.4 ms for 128
// columns x 10 M rows) from 96 columns on.  T4 = number of
// 4-column groups the chunk holds (even for TS = 8), T = number of TS-column tiles.
template <int TS>
__global__ void __launch_bounds__(256, 1)
pairwise_gap_kernel(const int32_t* __restrict__ bins, int64_t ldb, int64_t n, int d, int T4, int R,
                    int P, const int2* __restrict__ pair_tiles, int n_pair_tiles,
                    int32_t* __restrict__ gap) {
    extern __shared__ float chunk[];
    const int stride = 4 * (T4 + 1 + (T4 & 1));   // odd number of float4 per row
    const int slot = threadIdx.x % P, slice = threadIdx.x / P, S = 256 / P;
    const int pt = (int)blockIdx.y + slot * (int)gridDim.y;
    const bool active = pt < n_pair_tiles;
    const int2 t = active ? pair_tiles[pt] : make_int2(0, 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float m[TS][TS];
#pragma unroll
    for (int x = 0; x < TS; ++x)
#pragma unroll
        for (int y = 0; y < TS; ++y) m[x][y] = 0.f;

    const int64_t n_chunks = (n + R - 1) / R;
    const int patches_r = R / 8, n_patches = patches_r * T4;
    for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        const int64_t r0 = ch * R;
        __syncthreads();
        // load: a warp moves an 8-row x 4-column patch per step (32-byte global segments,
        // conflict-free shared-memory stores)
        for (int q = warp; q < n_patches; q += 8) {
            const int pc = q / patches_r, pr = q - pc * patches_r;
            const int r = pr * 8 + (lane & 7), c = pc * 4 + (lane >> 3);
            const int64_t row = r0 + r;
            chunk[r * stride + c] = (c < d && row < n) ? (float)bins[(int64_t)c * ldb + row] : 0.f;
        }
        __syncthreads();
        if (active) {
            const float* pa = chunk + TS * t.x;
            const float* pb = chunk + TS * t.y;
#pragma unroll 2
            for (int r = slice; r < R; r += S) {
                float av[TS], bv[TS];
#pragma unroll
                for (int v = 0; v < TS / 4; ++v) {
                    const float4 a = *reinterpret_cast<const float4*>(pa + r * stride + 4 * v);
                    const float4 b = *reinterpret_cast<const float4*>(pb + r * stride + 4 * v);
                    av[4 * v] = a.x; av[4 * v + 1] = a.y; av[4 * v + 2] = a.z; av[4 * v + 3] = a.w;
                    bv[4 * v] = b.x; bv[4 * v + 1] = b.y; bv[4 * v + 2] = b.z; bv[4 * v + 3] = b.w;
                }
#pragma unroll
                for (int x = 0; x < TS; ++x)
#pragma unroll
                    for (int y = 0; y < TS; ++y) m[x][y] = fmaxf(m[x][y], fabsf(av[x] - bv[y]));
            }
        }
    }
    if (active) {
#pragma unroll
        for (int x = 0; x < TS; ++x)
#pragma unroll
            for (int y = 0; y < TS; ++y) {
                const int i = TS * t.x + x, j = TS * t.y + y;
                if (i < d && j < d && i != j && m[x][y] > 0.f) {
                    atomicMax(&gap[i * d + j], (int)m[x][y]);
                    atomicMax(&gap[j * d + i], (int)m[x][y]);
                }
            }
    }
}

template <typename T>
int bin_columns(gr_pruner* h, const T* X, int64_t ldx, int32_t d, double frac, int32_t* bins,
                int64_t ldb, cudaStream_t st) {
    constexpr int G = kGroupBytes / (int)sizeof(T);
    const int64_t n = h->n;
    T* colbuf = static_cast<T*>(h->colbuf);
    T* keys_sorted = static_cast<T*>(h->keys_sorted);
    T* tops = static_cast<T*>(h->tops);
    const unsigned blocks_n = (unsigned)ceil_div<int64_t>(n, 256);
    for (int c0 = 0; c0 < d; c0 += G) {
        const int cols = std::min(G, d - c0);
        extract_columns_kernel<T><<<(unsigned)ceil_div<int64_t>(n, 32), 256, 0, st>>>(
            X, ldx, n, c0, cols, colbuf);
        count_launch();
        for (int c = 0; c < cols; ++c) {
            size_t temp = h->cub_temp_bytes;
            GR_CUDA_TRY(cub::DeviceRadixSort::SortKeys(h->cub_temp, temp, colbuf + (int64_t)c * n,
                                                       keys_sorted, (int)n, 0, (int)sizeof(T) * 8,
                                                       st));
            bin_bounds_kernel<T><<<1, 32, 0, st>>>(keys_sorted, n, frac, h->bounds, tops,
                                                   h->n_bounds);
            assign_bins_kernel<T><<<blocks_n, 256, 0, st>>>(colbuf + (int64_t)c * n, tops,
                                                            h->n_bounds, n,
                                                            bins + (int64_t)(c0 + c) * ldb);
            count_launch(2);
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return fail(GR_ERR_CUDA, "binning kernels failed to launch: %s", cudaGetErrorString(e));
    return GR_OK;
}

}  // namespace

extern "C" int gr_pruner_create(gr_pruner_t** out, int64_t n_rows, int device) {
    GR_REQUIRE(out != nullptr, "gr_pruner_create: out is NULL");
    *out = nullptr;
    GR_REQUIRE(n_rows >= 1 && n_rows < ((int64_t)1 << 31) - 64,
               "gr_pruner_create: n_rows = %lld (1 .. 2^31 - 65)", (long long)n_rows);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "gr_pruner_create: cannot select device %d", device);
    if (int rc = require_sm100(device)) return rc;
    gr_pruner* h = new (std::nothrow) gr_pruner();
    if (!h) return fail(GR_ERR_OUT_OF_MEMORY, "gr_pruner_create: host allocation failed");
    h->device = device;
    h->n = n_rows;
    cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&h->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    size_t t32 = 0, t64 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, t32, (const float*)nullptr, (float*)nullptr,
                                   (int)n_rows);
    cub::DeviceRadixSort::SortKeys(nullptr, t64, (const double*)nullptr, (double*)nullptr,
                                   (int)n_rows);
    h->cub_temp_bytes = std::max(t32, t64);
    const size_t n = (size_t)n_rows;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(p, bytes);
    };
    alloc(&h->colbuf, n * kGroupBytes);
    alloc(&h->keys_sorted, n * 8);
    alloc(&h->tops, (n + 1) * 8);
    alloc((void**)&h->bounds, (n + 1) * 4);
    alloc((void**)&h->n_bounds, 4);
    alloc(&h->cub_temp, std::max<size_t>(h->cub_temp_bytes, 16));
    if (e != cudaSuccess) {
        cudaGetLastError();
        gr_pruner_destroy(h);
        return fail(e == cudaErrorMemoryAllocation ? GR_ERR_OUT_OF_MEMORY : GR_ERR_CUDA,
                    "gr_pruner_create: workspace allocation failed: %s", cudaGetErrorString(e));
    }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        gr_pruner_destroy(h);
        return fail(GR_ERR_CUDA, "gr_pruner_create: %s", cudaGetErrorString(e));
    }
    *out = h;
    return GR_OK;
}

extern "C" int gr_pruner_destroy(gr_pruner_t* h) {
    if (!h) return GR_OK;
    DeviceGuard guard(h->device);
    cudaFree(h->colbuf);
    cudaFree(h->keys_sorted);
    cudaFree(h->tops);
    cudaFree(h->bounds);
    cudaFree(h->n_bounds);
    cudaFree(h->cub_temp);
    cudaFree(h->pair_tiles);
    delete h;
    return GR_OK;
}

static int check_bin_args(const gr_pruner* h, const void* X, int64_t ldx, int32_t d, double frac,
                          const int32_t* bins, int64_t ldb) {
    GR_REQUIRE(h != nullptr, "gr_prune_bin: pruner is NULL");
    GR_REQUIRE(d >= 0, "gr_prune_bin: d = %d", d);
    if (d == 0) return GR_OK;
    GR_REQUIRE(X != nullptr && bins != nullptr, "gr_prune_bin: NULL matrix");
    GR_REQUIRE(ldx >= d && ldb >= h->n, "gr_prune_bin: ldx = %lld < d or ldb = %lld < n_rows",
               (long long)ldx, (long long)ldb);
    // same refusal as prune.py:19-20
    GR_REQUIRE(frac > 0.0 && frac < 1.0, "must specify frac in interval (0, 1)");
    return GR_OK;
}

extern "C" int gr_prune_bin_f32(gr_pruner_t* h, const float* X_dev, int64_t ldx, int32_t d,
                                double frac, int32_t* bins_dev, int64_t ldb, void* stream) {
    if (int rc = check_bin_args(h, X_dev, ldx, d, frac, bins_dev, ldb)) return rc;
    if (d == 0) return GR_OK;
    DeviceGuard guard(h->device);
    return bin_columns<float>(h, X_dev, ldx, d, frac, bins_dev, ldb,
                              static_cast<cudaStream_t>(stream));
}

extern "C" int gr_prune_bin_f64(gr_pruner_t* h, const double* X_dev, int64_t ldx, int32_t d,
                                double frac, int32_t* bins_dev, int64_t ldb, void* stream) {
    if (int rc = check_bin_args(h, X_dev, ldx, d, frac, bins_dev, ldb)) return rc;
    if (d == 0) return GR_OK;
    DeviceGuard guard(h->device);
    return bin_columns<double>(h, X_dev, ldx, d, frac, bins_dev, ldb,
                               static_cast<cudaStream_t>(stream));
}

extern "C" int gr_prune_pairwise_gap_i32(gr_pruner_t* h, const int32_t* bins_dev, int64_t ldb,
                                         int32_t d, int32_t* gap_dev, void* stream) {
    GR_REQUIRE(h != nullptr, "gr_prune_pairwise_gap_i32: pruner is NULL");
    GR_REQUIRE(d >= 0 && d <= 1024, "gr_prune_pairwise_gap_i32: d = %d (0 .. 1024)", d);
    if (d == 0) return GR_OK;
    GR_REQUIRE(bins_dev != nullptr && gap_dev != nullptr && ldb >= h->n,
               "gr_prune_pairwise_gap_i32: NULL matrix or ldb < n_rows");
    DeviceGuard guard(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GR_CUDA_TRY(cudaMemsetAsync(gap_dev, 0, (size_t)d * d * sizeof(int32_t), st));
    if (d == 1) return GR_OK;
    // 8 x 8 register tiles once there are enough pair tiles to fill the slots, 4 x 4 below
    // (measured, 10 M rows: 64 columns 4.1 ms with 4 x 4 vs 6.6 ms with 8 x 8; 128 columns 29.9 vs
    // 18.4 ms; 256 columns 137.6 vs 74.2 ms -- profiles/r2_pairwise_tiles_integrated.txt)
    const int TS = (d >= 96 && !getenv("GR_PRUNE_TILE4")) ? 8 : 4;
    const int T = ceil_div(d, TS);
    const int T4 = T * (TS / 4);
    const int n_pair_tiles = T * (T + 1) / 2;
    if (h->pair_tiles_for_d != T) {
        // the table buffer is sized once for 256 tiles per side: re-filling it is a stream-ordered
        // kernel, no allocation and no synchronisation per call (d changes with every generation)
        if (!h->pair_tiles) {
            constexpr int kMaxT = 256;
            GR_CUDA_TRY(cudaMalloc((void**)&h->pair_tiles,
                                   (size_t)(kMaxT * (kMaxT + 1) / 2) * sizeof(int2)));
        }
        h->pair_tiles_for_d = T;
        pair_tile_table_kernel<<<dim3(T, ceil_div(T, 64)), 64, 0, st>>>(h->pair_tiles, T);
        count_launch();
    }
    // rows per shared-memory chunk: as many as fit (multiple of 8, at most 256)
    const int stride = 4 * (T4 + 1 + (T4 & 1));
    const int budget = h->max_smem - 1024;
    int R = std::min(256, budget / (stride * 4));
    R -= R % 8;
    GR_REQUIRE(R >= 8, "gr_prune_pairwise_gap_i32: d = %d does not fit shared memory", d);
    const int64_t n_chunks = ceil_div<int64_t>(h->n, R);
    const size_t smem = (size_t)R * stride * sizeof(float);
    int P = 1;
    while (P < 256 && P < n_pair_tiles) P *= 2;
    const int batches = ceil_div(n_pair_tiles, P);
    // one CTA per SM when the chunk takes most of the shared memory, more when it is small
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (size_t)budget / (smem + 1024)));
    const int workers = (int)std::min<int64_t>(n_chunks, std::max(1, per_sm * h->sms / batches));
    if (TS == 8) {
        GR_CUDA_TRY(cudaFuncSetAttribute(pairwise_gap_kernel<8>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pairwise_gap_kernel<8><<<dim3(workers, batches), 256, smem, st>>>(
            bins_dev, ldb, h->n, d, T4, R, P, h->pair_tiles, n_pair_tiles, gap_dev);
    } else {
        GR_CUDA_TRY(cudaFuncSetAttribute(pairwise_gap_kernel<4>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pairwise_gap_kernel<4><<<dim3(workers, batches), 256, smem, st>>>(
            bins_dev, ldb, h->n, d, T4, R, P, h->pair_tiles, n_pair_tiles, gap_dev);
    }
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return fail(GR_ERR_CUDA, "pairwise_gap_kernel launch failed: %s", cudaGetErrorString(e));
    return GR_OK;
}
