// Peer-visible device buffers and a flag barrier for the node-range sharded recursion
// (one process per GPU, all GPUs on one NVSwitch box).
//
// The exchange step of a sharded ReFeX level (every rank needs all rows
// of the previous level) is fused into the gather kernel: refex_gather_bcast_kernel stores each
// mean row into every rank's replica of the next input matrix, the remote ones through
// NVLink-mapped pointers.  That needs (1) buffers another process can map -- cudaMalloc +
// cudaIpcGetMemHandle / cudaIpcOpenMemHandle -- and (2) a stream-ordered barrier between levels,
// done here with system-scope flags in the same mapped memory instead of a collective library
// call (one single-warp kernel, a few microseconds over NVSwitch).

#include "common.cuh"

using namespace gr;

static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(gr_ipc_handle_t),
              "gr_ipc_handle_t must have the size of cudaIpcMemHandle_t");

extern "C" int gr_peer_alloc(void** dev_ptr_out, int64_t bytes, int device,
                             gr_ipc_handle_t* handle_out) {
    GR_REQUIRE(dev_ptr_out != nullptr && handle_out != nullptr, "gr_peer_alloc: NULL argument");
    GR_REQUIRE(bytes > 0, "gr_peer_alloc: bytes = %lld", (long long)bytes);
    *dev_ptr_out = nullptr;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "gr_peer_alloc: cannot select device %d", device);
    if (int rc = require_sm100(device)) return rc;
    void* p = nullptr;
    GR_CUDA_TRY(cudaMalloc(&p, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        cudaGetLastError();
        return fail(GR_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    memcpy(handle_out->bytes, &h, sizeof(h));
    *dev_ptr_out = p;
    return GR_OK;
}

extern "C" int gr_peer_open(void** dev_ptr_out, const gr_ipc_handle_t* handle, int device) {
    GR_REQUIRE(dev_ptr_out != nullptr && handle != nullptr, "gr_peer_open: NULL argument");
    *dev_ptr_out = nullptr;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "gr_peer_open: cannot select device %d", device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle->bytes, sizeof(h));
    void* p = nullptr;
    GR_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr_out = p;
    return GR_OK;
}

extern "C" int gr_peer_close(void* dev_ptr, int device) {
    if (!dev_ptr) return GR_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "gr_peer_close: cannot select device %d", device);
    GR_CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
    return GR_OK;
}

extern "C" int gr_peer_free(void* dev_ptr, int device) {
    if (!dev_ptr) return GR_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "gr_peer_free: cannot select device %d", device);
    GR_CUDA_TRY(cudaFree(dev_ptr));
    return GR_OK;
}

namespace {

constexpr int kMaxRanks = 16;

struct FlagPtrs {
    unsigned long long* flags[kMaxRanks];   // flags[q] = rank q's flag array (kMaxRanks + 1 words)
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Thread q < n_ranks: publish `epoch` into slot [rank] of rank q's flag array, then wait until
// slot [q] of the own array reaches `epoch`.  Everything this stream did before (the gather
// kernel's peer stores completed with that kernel) is ordered before the release store; the
// acquire loads order the peers' stores before whatever the stream runs next.  A rank that
// never arrives trips the timeout instead of hanging the GPU: slot [kMaxRanks] of the own array
// is set and the host reports it (gr_peer_barrier_status).
__global__ void peer_barrier_kernel(const FlagPtrs fp, int n_ranks, int rank,
                                    unsigned long long epoch, unsigned long long timeout_ns) {
    const int q = threadIdx.x;
    if (q >= n_ranks) return;
    __threadfence_system();
    st_release_sys(fp.flags[q] + rank, epoch);
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(fp.flags[rank] + q) < epoch) {
        if (globaltimer_ns() - t0 > timeout_ns) {
            st_release_sys(fp.flags[rank] + kMaxRanks, epoch);
            return;
        }
        __nanosleep(200);
    }
}

}  // namespace

extern "C" int gr_peer_barrier(void* const* flag_arrays, int32_t n_ranks, int32_t rank,
                               int64_t epoch, double timeout_s, void* stream) {
    GR_REQUIRE(flag_arrays != nullptr, "gr_peer_barrier: flag_arrays is NULL");
    GR_REQUIRE(n_ranks >= 1 && n_ranks <= kMaxRanks && rank >= 0 && rank < n_ranks,
               "gr_peer_barrier: rank %d of %d (at most %d ranks)", rank, n_ranks, kMaxRanks);
    GR_REQUIRE(epoch > 0, "gr_peer_barrier: epochs start at 1 and must increase");
    FlagPtrs fp;
    for (int q = 0; q < kMaxRanks; ++q) {
        fp.flags[q] = q < n_ranks ? static_cast<unsigned long long*>(flag_arrays[q]) : nullptr;
        GR_REQUIRE(q >= n_ranks || fp.flags[q] != nullptr, "gr_peer_barrier: flag array %d is NULL", q);
    }
    const double t = timeout_s > 0 ? timeout_s : 20.0;
    peer_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
        fp, n_ranks, rank, (unsigned long long)epoch, (unsigned long long)(t * 1e9));
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return fail(GR_ERR_CUDA, "peer_barrier_kernel launch failed: %s", cudaGetErrorString(e));
    return GR_OK;
}

extern "C" int64_t gr_peer_flag_words(void) { return kMaxRanks + 1; }

extern "C" int gr_peer_barrier_status(const void* own_flag_array, int64_t* timed_out_epoch) {
    GR_REQUIRE(own_flag_array != nullptr && timed_out_epoch != nullptr,
               "gr_peer_barrier_status: NULL argument");
    unsigned long long v = 0;
    GR_CUDA_TRY(cudaMemcpy(&v, static_cast<const unsigned long long*>(own_flag_array) + kMaxRanks,
                           sizeof(v), cudaMemcpyDeviceToHost));
    *timed_out_epoch = (int64_t)v;
    return GR_OK;
}
