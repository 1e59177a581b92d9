// Path B, tensor-core path (tcgen05 + TMA).  Under construction in this round: until the fused
// kernel lands, no shape is reported as supported and gr_nmf_mu_f32 runs the FFMA kernels.
#include "nmf_handle.cuh"

bool gr::nmf_tc_supported(const gr_nmf*, const float*, int64_t) { return false; }
int gr::nmf_iteration_tc(gr_nmf*, const float*, int64_t, float*, float*, cudaStream_t) {
    return gr::fail(GR_ERR_CUDA, "tcgen05 NMF path not built");
}
void gr::nmf_tc_release(gr_nmf*) {}
