// Round-2 experiment (not built by build(), not part of the library): can the sharded recursion's
// exchange use NVSwitch multicast?  With peer stores every rank writes each of its mean rows to
// world-1 replicas (measured ~0.6 TB/s of NVLink egress, the bound at 4-8 GPUs); with a multicast
// mapping (cuMulticastCreate + multimem.st) a rank stores the row ONCE and the switch replicates
// it.  This single-process probe answers, on the box it runs on:
//   1. do the devices report multicast / fabric / posix-fd handle support;
//   2. does a multicast object over all devices bind and map;
//   3. does one multimem.st from device 0 land in every device's memory, and how fast is a
//      streaming multimem.st of `bytes` compared with world-1 plain peer stores.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp_nvls_probe tools/exp_nvls_probe.cu -lcuda
// Run:   ./tools/exp_nvls_probe [MiB per device, default 256]
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <vector>

#define CU(x)                                                                         \
    do {                                                                              \
        CUresult r_ = (x);                                                            \
        if (r_ != CUDA_SUCCESS) {                                                     \
            const char* s_ = nullptr;                                                 \
            cuGetErrorString(r_, &s_);                                                \
            printf("FAIL %s -> %d (%s) at line %d\n", #x, (int)r_, s_ ? s_ : "?", __LINE__); \
            return 1;                                                                 \
        }                                                                             \
    } while (0)
#define RT(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            printf("FAIL %s -> %s at line %d\n", #x, cudaGetErrorString(e_), __LINE__); \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

// every thread stores 16 bytes through the multicast address: one store, all replicas
__global__ void mc_store_kernel(float4* mc, size_t n_vec, float v) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                     ::"l"(mc + i), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
    }
}
// the same bytes as plain stores to each replica's unicast address (what round 1 does)
struct Ptrs { float4* p[8]; int n; };
__global__ void uc_store_kernel(Ptrs ptrs, size_t n_vec, float v) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float4 val = make_float4(v, v, v, v);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride)
        for (int q = 0; q < ptrs.n; ++q) ptrs.p[q][i] = val;
}

int main(int argc, char** argv) {
    const size_t mib = argc > 1 ? (size_t)atoi(argv[1]) : 256;
    CU(cuInit(0));
    int n_dev = 0;
    CU(cuDeviceGetCount(&n_dev));
    if (n_dev > 8) n_dev = 8;
    printf("devices: %d\n", n_dev);
    std::vector<CUdevice> dev(n_dev);
    bool all_mc = true;
    for (int d = 0; d < n_dev; ++d) {
        CU(cuDeviceGet(&dev[d], d));
        int mc = 0, fabric = 0, posix = 0, vmm = 0;
        cuDeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev[d]);
        cuDeviceGetAttribute(&fabric, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, dev[d]);
        cuDeviceGetAttribute(&posix, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED,
                             dev[d]);
        cuDeviceGetAttribute(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev[d]);
        printf("  device %d: multicast %d, fabric handles %d, posix fd handles %d, vmm %d\n", d, mc,
               fabric, posix, vmm);
        all_mc = all_mc && mc;
    }
    if (n_dev < 2 || !all_mc) {
        printf("RESULT: multicast not available on this box (needs >= 2 devices reporting it)\n");
        return 0;
    }
    // primary contexts through the runtime so kernels can be launched with <<<>>>
    for (int d = 0; d < n_dev; ++d) { RT(cudaSetDevice(d)); RT(cudaFree(0)); }
    for (int d = 0; d < n_dev; ++d)
        for (int q = 0; q < n_dev; ++q)
            if (q != d) { cudaSetDevice(d); cudaDeviceEnablePeerAccess(q, 0); cudaGetLastError(); }

    CUmulticastObjectProp mcp = {};
    mcp.numDevices = (unsigned)n_dev;
    mcp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    mcp.flags = 0;
    size_t gran = 0;
    mcp.size = mib << 20;
    CU(cuMulticastGetGranularity(&gran, &mcp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    const size_t size = ((mib << 20) + gran - 1) / gran * gran;
    mcp.size = size;
    printf("multicast granularity %zu bytes, object size %zu bytes\n", gran, size);
    CUmemGenericAllocationHandle mc_handle;
    CU(cuMulticastCreate(&mc_handle, &mcp));
    for (int d = 0; d < n_dev; ++d) CU(cuMulticastAddDevice(mc_handle, dev[d]));

    std::vector<CUmemGenericAllocationHandle> mem(n_dev);
    std::vector<CUdeviceptr> uc(n_dev), mcva(n_dev);
    for (int d = 0; d < n_dev; ++d) {
        RT(cudaSetDevice(d));
        CUmemAllocationProp prop = {};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = d;
        prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        CU(cuMemCreate(&mem[d], size, &prop, 0));
        CU(cuMulticastBindMem(mc_handle, 0, mem[d], 0, size, 0));
        // unicast mapping of the device's own physical memory, readable by every device
        CU(cuMemAddressReserve(&uc[d], size, gran, 0, 0));
        CU(cuMemMap(uc[d], size, 0, mem[d], 0));
        std::vector<CUmemAccessDesc> acc(n_dev);
        for (int q = 0; q < n_dev; ++q) {
            acc[q].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
            acc[q].location.id = q;
            acc[q].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        }
        CU(cuMemSetAccess(uc[d], size, acc.data(), (size_t)n_dev));
    }
    for (int d = 0; d < n_dev; ++d) {
        // multicast mapping as seen from device d
        RT(cudaSetDevice(d));
        CU(cuMemAddressReserve(&mcva[d], size, gran, 0, 0));
        CU(cuMemMap(mcva[d], size, 0, mc_handle, 0));
        CUmemAccessDesc acc = {};
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = d;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        CU(cuMemSetAccess(mcva[d], size, &acc, 1));
    }
    printf("multicast object bound and mapped on all devices\n");

    // ---- correctness: one multimem.st pass from device 0 reaches every replica ----
    const size_t n_vec = size / sizeof(float4);
    for (int d = 0; d < n_dev; ++d) { RT(cudaSetDevice(d)); RT(cudaMemset((void*)uc[d], 0, size)); }
    for (int d = 0; d < n_dev; ++d) { RT(cudaSetDevice(d)); RT(cudaDeviceSynchronize()); }
    RT(cudaSetDevice(0));
    mc_store_kernel<<<148 * 8, 256>>>((float4*)mcva[0], n_vec, 3.0f);
    RT(cudaDeviceSynchronize());
    bool ok = true;
    for (int d = 0; d < n_dev; ++d) {
        RT(cudaSetDevice(d));
        float first = 0, last = 0;
        RT(cudaMemcpy(&first, (void*)uc[d], 4, cudaMemcpyDeviceToHost));
        RT(cudaMemcpy(&last, (void*)(uc[d] + size - 4), 4, cudaMemcpyDeviceToHost));
        printf("  replica on device %d: first %.1f last %.1f\n", d, first, last);
        ok = ok && first == 3.0f && last == 3.0f;
    }
    printf("multimem.st reaches every replica: %s\n", ok ? "yes" : "NO");

    // ---- speed: multicast stores vs world-1 + 1 unicast stores of the same rows ----
    RT(cudaSetDevice(0));
    cudaEvent_t e0, e1;
    RT(cudaEventCreate(&e0));
    RT(cudaEventCreate(&e1));
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        RT(cudaEventRecord(e0));
        mc_store_kernel<<<148 * 8, 256>>>((float4*)mcva[0], n_vec, 4.0f);
        RT(cudaEventRecord(e1));
        RT(cudaEventSynchronize(e1));
        RT(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("multimem.st  : %zu MiB to %d replicas in %.3f ms = %.1f GB/s of payload\n", size >> 20,
           n_dev, ms, (double)size / ms / 1e6);
    Ptrs ptrs;
    ptrs.n = n_dev;
    for (int d = 0; d < n_dev; ++d) ptrs.p[d] = (float4*)uc[d];
    for (int rep = 0; rep < 2; ++rep) {
        RT(cudaEventRecord(e0));
        uc_store_kernel<<<148 * 8, 256>>>(ptrs, n_vec, 5.0f);
        RT(cudaEventRecord(e1));
        RT(cudaEventSynchronize(e1));
        RT(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("peer stores  : %zu MiB to %d replicas in %.3f ms = %.1f GB/s of payload (%.1f GB/s on the wire)\n",
           size >> 20, n_dev, ms, (double)size / ms / 1e6, (double)size * (n_dev - 1) / ms / 1e6);

    // ---- the exchange itself: EVERY device multicasts its own 1/n slice at the same time (an
    // all-gather by multicast); what bounds it is each device's ingress of (n-1)/n of the object
    {
        const size_t slice_vec = n_vec / n_dev;
        std::vector<cudaStream_t> st(n_dev);
        for (int d = 0; d < n_dev; ++d) { RT(cudaSetDevice(d)); RT(cudaStreamCreate(&st[d])); }
        double best_mc = 1e30, best_uc = 1e30;
        for (int rep = 0; rep < 3; ++rep) {
            for (int mode = 0; mode < 2; ++mode) {
                for (int d = 0; d < n_dev; ++d) { RT(cudaSetDevice(d)); RT(cudaDeviceSynchronize()); }
                timespec t0, t1;
                clock_gettime(CLOCK_MONOTONIC, &t0);
                for (int d = 0; d < n_dev; ++d) {
                    RT(cudaSetDevice(d));
                    if (mode == 0) {
                        mc_store_kernel<<<148 * 8, 256, 0, st[d]>>>(
                            (float4*)mcva[d] + (size_t)d * slice_vec, slice_vec, 6.0f);
                    } else {
                        Ptrs pp;
                        pp.n = n_dev;
                        for (int q = 0; q < n_dev; ++q) pp.p[q] = (float4*)uc[q] + (size_t)d * slice_vec;
                        uc_store_kernel<<<148 * 8, 256, 0, st[d]>>>(pp, slice_vec, 7.0f);
                    }
                }
                for (int d = 0; d < n_dev; ++d) { RT(cudaSetDevice(d)); RT(cudaStreamSynchronize(st[d])); }
                clock_gettime(CLOCK_MONOTONIC, &t1);
                const double dt = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) / 1e6;
                if (mode == 0) best_mc = dt < best_mc ? dt : best_mc;
                else best_uc = dt < best_uc ? dt : best_uc;
            }
        }
        const double slice_bytes = (double)slice_vec * sizeof(float4);
        const double ingress = slice_bytes * (n_dev - 1);
        printf("all devices at once, %d slices of %.1f MiB (host-timed, best of 3):\n", n_dev,
               slice_bytes / 1048576.0);
        printf("  multimem.st : %.3f ms = %.1f GB/s ingress per device, %.1f GB/s egress per device\n",
               best_mc, ingress / best_mc / 1e6, slice_bytes / best_mc / 1e6);
        printf("  peer stores : %.3f ms = %.1f GB/s ingress per device, %.1f GB/s egress per device\n",
               best_uc, ingress / best_uc / 1e6, ingress / best_uc / 1e6);
    }
    printf("RESULT: done\n");
    return 0;
}
