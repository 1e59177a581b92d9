// Library-level entry points and the CSR graph handle (include/graphrole_b200.h).
//
// gr_csr_create is the native form of the reference's graph plugin calls
// graph.get_nodes()/graph.get_neighbors(node) (graphrole/graph/interface/base.py:58-69,
// interface/networkx.py:36-46): the host side (graphrole_b200/graph) flattens them once into
// rowptr/colidx and every recursion level reuses the handle.

#include <algorithm>

#include <cstdlib>

#include "csr_handle.cuh"

using namespace gr;

extern "C" const char* gr_last_error(void) { return tls_error_buffer(); }

extern "C" const char* gr_version(void) { return "graphrole_b200 0.1.0 sm_100a"; }

extern "C" int64_t gr_kernel_launch_count(void) {
    return launch_counter().load(std::memory_order_relaxed);
}

namespace {

// Range check of colidx: counts entries outside [0, n_cols).
__global__ void __launch_bounds__(256)
colidx_range_kernel(const int32_t* __restrict__ colidx, int64_t nnz, int64_t n_cols,
                    unsigned long long* __restrict__ bad) {
    unsigned long long local = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) {
        const int32_t c = __ldcs(colidx + i);
        local += (c < 0 || (int64_t)c >= n_cols) ? 1ull : 0ull;
    }
    local = __reduce_add_sync(0xffffffffu, (unsigned)local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(bad, local);
}

// ---- GR_CSR_HOT_HINTS: in-degree -> threshold -> sign-bit tagged copy of colidx -------------
__global__ void __launch_bounds__(256)
indegree_kernel(const int32_t* __restrict__ colidx, int64_t nnz, int64_t n_cols,
                int32_t* __restrict__ indeg) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) {
        const int32_t c = __ldcs(colidx + i);
        if (c >= 0 && (int64_t)c < n_cols) atomicAdd(indeg + c, 1);   // unvalidated input stays safe
    }
}

constexpr int kDegBins = 1 << 16;

__global__ void __launch_bounds__(256)
degree_histogram_kernel(const int32_t* __restrict__ indeg, int64_t n,
                        unsigned long long* __restrict__ hist) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        atomicAdd(hist + min(indeg[i], kDegBins - 1), 1ull);
}

__global__ void __launch_bounds__(256)
tag_hot_kernel(const int32_t* __restrict__ colidx, int64_t nnz, int64_t n_cols,
               const int32_t* __restrict__ indeg, int32_t threshold,
               int32_t* __restrict__ tagged) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) {
        const int32_t c = __ldcs(colidx + i);
        const bool hot = c >= 0 && (int64_t)c < n_cols && __ldg(indeg + c) >= threshold;
        tagged[i] = hot ? (c | (int32_t)0x80000000) : c;
    }
}

// Tags arcs into the (at most) kHotBudgetBytes / kHotRowBytes rows of highest in-degree.
int build_hot_tags(gr_csr* g) {
    if (g->nnz == 0 || g->n_cols == 0) return GR_OK;
    int64_t budget_bytes = kHotBudgetBytes;
    if (const char* e = getenv("GR_CSR_HOT_BUDGET_MB")) budget_bytes = (int64_t)atoi(e) << 20;
    const int64_t budget_rows = std::min<int64_t>(budget_bytes / g->hot_row_bytes, g->n_cols);
    int32_t* indeg = nullptr;
    unsigned long long* hist = nullptr;
    std::vector<unsigned long long> h_hist(kDegBins);
    const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(g->nnz, 256), 148 * 32);
    int rc = GR_OK;
    auto cleanup = [&]() { cudaFree(indeg); cudaFree(hist); };
#define HOT_TRY(expr)                                                                       \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            cleanup();                                                                      \
            return fail(_e == cudaErrorMemoryAllocation ? GR_ERR_OUT_OF_MEMORY : GR_ERR_CUDA, \
                        "hot-row tagging: %s failed: %s", #expr, cudaGetErrorString(_e));   \
        }                                                                                   \
    } while (0)
    HOT_TRY(cudaMalloc(&indeg, (size_t)g->n_cols * sizeof(int32_t)));
    HOT_TRY(cudaMalloc(&hist, kDegBins * sizeof(unsigned long long)));
    HOT_TRY(cudaMemset(indeg, 0, (size_t)g->n_cols * sizeof(int32_t)));
    HOT_TRY(cudaMemset(hist, 0, kDegBins * sizeof(unsigned long long)));
    indegree_kernel<<<blocks, 256>>>(g->colidx, g->nnz, g->n_cols, indeg);
    degree_histogram_kernel<<<(int)std::min<int64_t>(ceil_div<int64_t>(g->n_cols, 256), 148 * 32),
                              256>>>(indeg, g->n_cols, hist);
    count_launch(2);
    HOT_TRY(cudaGetLastError());
    HOT_TRY(cudaMemcpy(h_hist.data(), hist, kDegBins * sizeof(unsigned long long),
                       cudaMemcpyDeviceToHost));
    // smallest threshold whose row count fits the budget (threshold >= 2: a row gathered once
    // has nothing to reuse)
    int64_t rows_at_or_above = 0;
    int32_t threshold = kDegBins;
    for (int b = kDegBins - 1; b >= 2; --b) {
        if (rows_at_or_above + (int64_t)h_hist[b] > budget_rows) break;
        rows_at_or_above += (int64_t)h_hist[b];
        threshold = b;
    }
    g->n_hot_rows = rows_at_or_above;
    if (rows_at_or_above > 0) {
        HOT_TRY(cudaMalloc(&g->d_colidx_tagged, (size_t)g->nnz * sizeof(int32_t)));
        tag_hot_kernel<<<blocks, 256>>>(g->colidx, g->nnz, g->n_cols, indeg, threshold,
                                         g->d_colidx_tagged);
        count_launch();
        HOT_TRY(cudaGetLastError());
        HOT_TRY(cudaDeviceSynchronize());
    }
#undef HOT_TRY
    cleanup();
    return rc;
}

template <typename T>
int upload(T** dst, const std::vector<T>& src) {
    *dst = nullptr;
    if (src.empty()) return GR_OK;
    GR_CUDA_TRY(cudaMalloc(dst, src.size() * sizeof(T)));
    GR_CUDA_TRY(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return GR_OK;
}

}  // namespace

extern "C" int gr_csr_create(gr_csr_t** out, int64_t n_rows, int64_t n_cols, int64_t nnz,
                             const int64_t* rowptr_dev, const int32_t* colidx_dev,
                             int device, int flags) {
    GR_REQUIRE(out != nullptr, "gr_csr_create: out is NULL");
    *out = nullptr;
    GR_REQUIRE(n_rows >= 0 && n_cols >= 0 && nnz >= 0, "gr_csr_create: negative size");
    GR_REQUIRE(n_cols <= (int64_t)INT32_MAX, "gr_csr_create: n_cols exceeds int32 colidx range");
    GR_REQUIRE(rowptr_dev != nullptr, "gr_csr_create: rowptr is NULL");
    GR_REQUIRE(nnz == 0 || colidx_dev != nullptr, "gr_csr_create: colidx is NULL with nnz > 0");
    GR_REQUIRE((reinterpret_cast<uintptr_t>(rowptr_dev) & 7u) == 0 &&
               (reinterpret_cast<uintptr_t>(colidx_dev) & 3u) == 0,
               "gr_csr_create: rowptr must be 8-byte and colidx 4-byte aligned");

    DeviceGuard guard(device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "gr_csr_create: cannot select device %d", device);
    if (int rc = require_sm100(device)) return rc;

    // rowptr is 8(n+1) bytes (80 MB at 10 M nodes): validate and find hub rows on the host.
    std::vector<int64_t> rp((size_t)n_rows + 1);
    GR_CUDA_TRY(cudaMemcpy(rp.data(), rowptr_dev, rp.size() * sizeof(int64_t),
                           cudaMemcpyDeviceToHost));
    // rowptr[0] > 0 is a row-range shard that keeps a few leading colidx entries of the previous
    // row so that its arcs sit at the same offsets modulo 32 as in the unsharded graph (the
    // gather kernel's chunking, hence its fp32 summation order, depends on them)
    if (rp[0] < 0 || rp[0] > nnz)
        return fail(GR_ERR_INVALID_GRAPH, "rowptr[0] = %lld, expected 0 <= rowptr[0] <= nnz",
                    (long long)rp[0]);
    if (rp[(size_t)n_rows] != nnz)
        return fail(GR_ERR_INVALID_GRAPH, "rowptr[n_rows] = %lld, expected nnz = %lld",
                    (long long)rp[(size_t)n_rows], (long long)nnz);

    gr_csr* g = new (std::nothrow) gr_csr();
    if (!g) return fail(GR_ERR_OUT_OF_MEMORY, "gr_csr_create: host allocation failed");
    g->device = device;
    g->n_rows = n_rows;
    g->n_cols = n_cols;
    g->nnz = nnz;
    g->rowptr = rowptr_dev;
    g->colidx = colidx_dev;

    if (const char* e = getenv("GR_REFEX_HUB_THRESHOLD")) g->hub_threshold = std::max(32, atoi(e));
    if (const char* e = getenv("GR_REFEX_HUB_SEGMENT")) g->hub_segment = std::max(32, atoi(e));
    std::vector<int64_t> seg_begin, seg_end;
    g->h_hub_seg_first.push_back(0);
    for (int64_t r = 0; r < n_rows; ++r) {
        const int64_t b = rp[(size_t)r], e = rp[(size_t)r + 1];
        if (e < b) {
            delete g;
            return fail(GR_ERR_INVALID_GRAPH, "rowptr decreases at row %lld (%lld -> %lld)",
                        (long long)r, (long long)b, (long long)e);
        }
        if (e - rp[(size_t)std::max<int64_t>(r - 31, 0)] >= (int64_t)1 << 31) {
            delete g;
            return fail(GR_ERR_INVALID_GRAPH, "rows %lld..%lld hold more than 2^31 arcs",
                        (long long)std::max<int64_t>(r - 31, 0), (long long)r);
        }
        if (e - b > g->hub_threshold) {
            g->h_hub_row.push_back(r);
            for (int64_t s = b; s < e; s += g->hub_segment) {
                seg_begin.push_back(s);
                seg_end.push_back(std::min(e, s + g->hub_segment));
            }
            g->h_hub_seg_first.push_back((int64_t)seg_begin.size());
        }
    }
    g->n_hub_rows = (int64_t)g->h_hub_row.size();
    g->n_segments = (int64_t)seg_begin.size();

    int rc = GR_OK;
    if ((rc = upload(&g->d_hub_row, g->h_hub_row)) ||
        (rc = upload(&g->d_hub_seg_first, g->h_hub_seg_first)) ||
        (rc = upload(&g->d_seg_begin, seg_begin)) || (rc = upload(&g->d_seg_end, seg_end))) {
        gr_csr_destroy(g);
        return rc;
    }

    if ((flags & GR_CSR_VALIDATE) && nnz > 0) {
        unsigned long long* d_bad = nullptr;
        unsigned long long h_bad = 0;
        cudaError_t e = cudaMalloc(&d_bad, sizeof(*d_bad));
        if (e == cudaSuccess) e = cudaMemset(d_bad, 0, sizeof(*d_bad));
        if (e == cudaSuccess) {
            const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(nnz, 256), 148 * 16);
            colidx_range_kernel<<<blocks, 256>>>(colidx_dev, nnz, n_cols, d_bad);
            count_launch();
            e = cudaGetLastError();
        }
        if (e == cudaSuccess)
            e = cudaMemcpy(&h_bad, d_bad, sizeof(h_bad), cudaMemcpyDeviceToHost);
        cudaFree(d_bad);
        if (e != cudaSuccess) {
            gr_csr_destroy(g);
            return fail(GR_ERR_CUDA, "colidx validation failed: %s", cudaGetErrorString(e));
        }
        if (h_bad) {
            gr_csr_destroy(g);
            return fail(GR_ERR_INVALID_GRAPH, "%llu colidx entries outside [0, %lld)", h_bad,
                        (long long)n_cols);
        }
    }

    if (flags & GR_CSR_HOT_HINTS) {
        if ((rc = build_hot_tags(g))) {
            gr_csr_destroy(g);
            return rc;
        }
    }

    *out = g;
    return GR_OK;
}

extern "C" int gr_csr_tune_hot_rows(gr_csr_t* g, int64_t row_bytes) {
    GR_REQUIRE(g != nullptr, "gr_csr_tune_hot_rows: handle is NULL");
    GR_REQUIRE(row_bytes >= 4, "gr_csr_tune_hot_rows: row_bytes = %lld", (long long)row_bytes);
    DeviceGuard guard(g->device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "cannot select device %d", g->device);
    // whole 32-byte sectors are what a gather occupies in L2
    const int64_t bytes = (row_bytes + 31) / 32 * 32;
    if (bytes == g->hot_row_bytes && g->d_colidx_tagged) return GR_OK;
    GR_CUDA_TRY(cudaDeviceSynchronize());
    cudaFree(g->d_colidx_tagged);
    g->d_colidx_tagged = nullptr;
    g->n_hot_rows = 0;
    g->hot_row_bytes = bytes;
    return build_hot_tags(g);
}

extern "C" int gr_csr_destroy(gr_csr_t* g) {
    if (!g) return GR_OK;
    DeviceGuard guard(g->device);
    cudaFree(g->d_hub_row);
    cudaFree(g->d_hub_seg_first);
    cudaFree(g->d_seg_begin);
    cudaFree(g->d_seg_end);
    cudaFree(g->d_partial);
    cudaFree(g->d_colidx_tagged);
    cudaFree(g->d_stage_x);
    cudaFree(g->d_stage_out[0]);
    cudaFree(g->d_stage_out[1]);
    if (g->copy_stream) cudaStreamDestroy(g->copy_stream);
    for (int i = 0; i < 2; ++i) {
        if (g->ev_level[i]) cudaEventDestroy(g->ev_level[i]);
        if (g->ev_copied[i]) cudaEventDestroy(g->ev_copied[i]);
    }
    delete g;
    return GR_OK;
}

extern "C" int gr_csr_info(const gr_csr_t* g, int64_t* n_rows, int64_t* n_cols, int64_t* nnz,
                           int64_t* n_hub_rows, int64_t* n_hub_segments, int64_t* n_hot_rows) {
    GR_REQUIRE(g != nullptr, "gr_csr_info: handle is NULL");
    if (n_rows) *n_rows = g->n_rows;
    if (n_cols) *n_cols = g->n_cols;
    if (nnz) *nnz = g->nnz;
    if (n_hub_rows) *n_hub_rows = g->n_hub_rows;
    if (n_hub_segments) *n_hub_segments = g->n_segments;
    if (n_hot_rows) *n_hot_rows = g->n_hot_rows;
    return GR_OK;
}
