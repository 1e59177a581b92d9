// gr_csr: the device-resident CSR graph handle and its hub-row decomposition.
#pragma once

#include <vector>

#include "common.cuh"

namespace gr {

// Rows with more arcs than kHubThreshold are not processed by a single warp: they are cut
// into segments of kHubSegment arcs, each reduced by its own warp into an fp32 partial, and a
// second kernel combines a row's partials in fp64 in segment order (deterministic, no atomics).
constexpr int64_t kHubThreshold = 2048;
constexpr int64_t kHubSegment = 1024;

}  // namespace gr

struct gr_csr {
    int device = 0;
    int64_t n_rows = 0, n_cols = 0, nnz = 0;
    const int64_t* rowptr = nullptr;  // caller-owned, device
    const int32_t* colidx = nullptr;  // caller-owned, device

    // hub decomposition (library-owned)
    int64_t n_hub_rows = 0, n_segments = 0;
    std::vector<int64_t> h_hub_row;        // [n_hub_rows] row id, ascending
    std::vector<int64_t> h_hub_seg_first;  // [n_hub_rows + 1] first segment of each hub row
    int64_t* d_hub_row = nullptr;
    int64_t* d_hub_seg_first = nullptr;
    int64_t* d_seg_begin = nullptr;  // [n_segments] arc range of each segment
    int64_t* d_seg_end = nullptr;

    // grow-only workspaces
    float* d_partial = nullptr;  // [n_segments, d] hub partial sums
    size_t partial_floats = 0;
    float* d_stage_x = nullptr;  // host-variant staging: input / recursion ping-pong
    size_t stage_x_floats = 0;
    float* d_stage_out[2] = {nullptr, nullptr};
    size_t stage_out_floats = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_level[2] = {nullptr, nullptr};
    cudaEvent_t ev_copied[2] = {nullptr, nullptr};
};
