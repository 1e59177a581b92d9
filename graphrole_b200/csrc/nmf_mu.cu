// Path B placeholder: replaced by the real kernels in the next milestone of this round.
#include "common.cuh"
using namespace gr;
struct gr_nmf { int device; };
extern "C" int gr_nmf_create(gr_nmf_t** out, int64_t, int32_t, int32_t, int) {
    if (out) *out = nullptr;
    return fail(GR_ERR_CUDA, "gr_nmf_*: kernels not built yet");
}
extern "C" int gr_nmf_destroy(gr_nmf_t*) { return GR_OK; }
extern "C" int gr_nmf_mu_f32(gr_nmf_t*, const float*, int64_t, float*, float*, int32_t, double,
                             int32_t, int32_t, int32_t*, double*, void*) {
    return fail(GR_ERR_CUDA, "gr_nmf_*: kernels not built yet");
}
extern "C" int gr_nmf_error_f32(gr_nmf_t*, const float*, int64_t, const float*, const float*,
                                double*, void*) {
    return fail(GR_ERR_CUDA, "gr_nmf_*: kernels not built yet");
}
