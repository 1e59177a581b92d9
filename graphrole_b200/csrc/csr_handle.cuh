// gr_csr: the device-resident CSR graph handle and its hub-row decomposition.
#pragma once

#include <vector>

#include "common.cuh"

namespace gr {

// Rows with more arcs than kHubThreshold are not processed by a single warp: they are cut
// into segments of kHubSegment arcs, each reduced by its own warp into an fp32 partial, and a
// second kernel combines a row's partials in fp64 in segment order (deterministic, no atomics).
// Defaults; GR_REFEX_HUB_THRESHOLD / GR_REFEX_HUB_SEGMENT override them when a handle is created
// (the split is a property of the handle: every shard of a graph must be created with the same
// values to reproduce the unsharded bits).
constexpr int64_t kHubThreshold = 2048;
constexpr int64_t kHubSegment = 1024;

// GR_CSR_HOT_HINTS: rows kept in L2 with an evict_last policy.  Measured on B200 (C3, d = 64):
// the level time is flat for 150 k - 400 k hot rows of 256 B and worse outside, i.e. about
// 0.4 of the 126 MB L2; the tag is computed for this byte budget at 256-byte rows.
constexpr int64_t kHotBudgetBytes = 48ll << 20;
constexpr int64_t kHotRowBytes = 256;

}  // namespace gr

struct gr_csr {
    int device = 0;
    int64_t n_rows = 0, n_cols = 0, nnz = 0;
    const int64_t* rowptr = nullptr;  // caller-owned, device
    const int32_t* colidx = nullptr;  // caller-owned, device
    int32_t* d_colidx_tagged = nullptr;  // library-owned copy, sign bit = "hot row" (optional)
    int64_t n_hot_rows = 0;
    int64_t hot_row_bytes = gr::kHotRowBytes;  // feature-row width the hot-row budget assumes

    // hub decomposition (library-owned)
    int64_t hub_threshold = gr::kHubThreshold, hub_segment = gr::kHubSegment;
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
