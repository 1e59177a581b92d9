// Shared host/device helpers for libgraphrole_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/graphrole_b200.h"

namespace gr {

// ---- error reporting ------------------------------------------------------------------
inline char* tls_error_buffer() {
    static thread_local char buf[1024] = {0};
    return buf;
}

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tls_error_buffer(), 1024, fmt, ap);
    va_end(ap);
    return code;
}

#define GR_CUDA_TRY(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            cudaGetLastError();                                                             \
            return gr::fail(_e == cudaErrorMemoryAllocation ? GR_ERR_OUT_OF_MEMORY          \
                                                            : GR_ERR_CUDA,                  \
                            "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,            \
                            cudaGetErrorString(_e));                                        \
        }                                                                                   \
    } while (0)

#define GR_REQUIRE(cond, ...)                                                               \
    do {                                                                                    \
        if (!(cond)) return gr::fail(GR_ERR_INVALID_ARGUMENT, __VA_ARGS__);                 \
    } while (0)

// ---- launch accounting (bench.py reports it as gpu_launches) ---------------------------
inline std::atomic<int64_t>& launch_counter() {
    static std::atomic<int64_t> c{0};
    return c;
}
inline void count_launch(int n = 1) { launch_counter().fetch_add(n, std::memory_order_relaxed); }

// Scoped device switch for entry points that take a device ordinal or a handle.
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Refuse to run anywhere but Blackwell datacenter parts: the .so carries only sm_100a SASS.
inline int require_sm100(int device) {
    cudaDeviceProp prop;
    GR_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(GR_ERR_UNSUPPORTED_DEVICE,
                    "device %d is sm_%d%d; libgraphrole_b200 is built for sm_100a only and has "
                    "no fallback path", device, prop.major, prop.minor);
    return GR_OK;
}

// Stream-ordered workspaces (cudaMallocAsync) come out of the device's default memory pool, whose
// release threshold is 0 by default: everything goes back to the OS at the next synchronisation
// and a 2 GB workspace is then re-created on every call (measured: 80 - 180 ms on top of a 62 ms
// level-0 pass).  Let the pool keep up to 4 GB between calls; done once per device.
inline void retain_async_pool(int device) {
    static std::atomic<unsigned> done{0};
    const unsigned bit = 1u << (device & 31);
    if (done.load(std::memory_order_relaxed) & bit) return;
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
        uint64_t threshold = 4ull << 30;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    cudaGetLastError();
    done.fetch_or(bit, std::memory_order_relaxed);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

}  // namespace gr
