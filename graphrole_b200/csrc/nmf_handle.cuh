// gr_nmf: workspaces of the NMF multiplicative-update loop and the internal entry points
// shared by the FFMA path (nmf_mu.cu) and the tcgen05 path (nmf_mu_tc.cu).
#pragma once

#include <vector>

#include "common.cuh"

struct gr_nmf {
    int device = 0;
    int64_t n = 0;
    int f = 0, r = 0;
    int rp = 0;                  // r padded to 8 / 16 / 32
    int splits = 0;              // row splits (= number of partial W^T X / W^T W blocks)
    int64_t rows_per_split = 0;
    float* d_hht = nullptr;      // [r, r]
    float* d_wtw = nullptr;      // [r, r]
    float* d_h_next = nullptr;   // [r, f] H update target (copied back into H)
    float* d_part_wtx = nullptr; // [splits, rp, f]
    float* d_part_wtw = nullptr; // [splits, rp, r]
    double* d_err_part = nullptr;
    std::vector<double> h_err_part;
    bool last_path_tc = false;
    void* tc_state = nullptr;    // owned by nmf_mu_tc.cu
    // r % 4 != 0 on the tensor-core path: rows of W [n, r] are not 16-byte multiples (TMA), so
    // the loop runs on zero-padded factors of rank r4 = r rounded up to 4 (nmf_mu.cu)
    // row-sharded runs (gr_nmf_iteration_local_f32): stop an iteration before the reduction of
    // the partial sums and remember where they are
    bool defer_finish = false;
    const float* pending_wtx = nullptr;
    const float* pending_wtw = nullptr;
    int pending_splits = 0, pending_rp = 0;
    gr_nmf* padded = nullptr;    // handle of rank r4
    float* d_wpad = nullptr;     // [n, r4]
    float* d_hpad = nullptr;     // [r4, f]
};

namespace gr {

int nmf_iteration_fma(gr_nmf* h, const float* X, int64_t ldx, float* W, float* H, cudaStream_t st);
int nmf_hht(gr_nmf* h, const float* H, cudaStream_t st);
// Shared tail of an iteration: reduce [splits, rp, *] partials in fixed order, update H.
int nmf_finish_iteration(gr_nmf* h, const float* part_wtx, const float* part_wtw, int splits,
                         int rp, float* H, cudaStream_t st);
int nmf_error(gr_nmf* h, const float* X, int64_t ldx, const float* W, const float* H, double* err,
              cudaStream_t st);

// tcgen05 / TMA fused single-pass iteration (nmf_mu_tc.cu)
bool nmf_tc_supported(const gr_nmf* h, const float* X, int64_t ldx);
int nmf_iteration_tc(gr_nmf* h, const float* X, int64_t ldx, float* W, float* H, cudaStream_t st);
// ||X - W H||_F with W H on the tensor core (TF32 operands, fp32 residual, fp64 sums)
int nmf_error_tc(gr_nmf* h, const float* X, int64_t ldx, const float* W, const float* H,
                 double* err, cudaStream_t st);
void nmf_tc_release(gr_nmf* h);

}  // namespace gr
