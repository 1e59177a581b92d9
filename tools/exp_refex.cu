// Development harness (not part of the product library): kernel variants of the ReFeX gather
// for A/B timing on the GPU box.  Built into tools/libexp_refex.so and driven by
// tools/exp_refex.py.  d is fixed per instantiation (64 -> LPR 16, 32 -> LPR 8), fp32, vec4.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int64_t kHub = 2048;

struct Args {
    const int64_t* __restrict__ rowptr;
    const int32_t* __restrict__ colidx;
    const float* __restrict__ X;
    int64_t ldx;
    int64_t n_rows;
    float* __restrict__ out;   // [n, 2d]: sum | mean
    int32_t rows_per_warp;
    int32_t hot_rows;          // idx < hot_rows -> L2 evict_last hint (HINT variants)
};

__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ld_hint(const float* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float4 ld_hint_na(const float* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float4 ld_noalloc(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// MODE: 0 plain __ldg, 1 L1::no_allocate, 2 hot rows evict_last / others normal,
//       3 hot rows evict_last / others evict_first
template <int MODE>
__device__ __forceinline__ float4 load_row(const float* p, bool ok, bool hot, uint64_t pl,
                                           uint64_t pf) {
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!ok) return z;
    if (MODE == 0) return __ldg(reinterpret_cast<const float4*>(p));
    if (MODE == 1) return ld_noalloc(p);
    if (MODE == 2) return hot ? ld_hint(p, pl) : __ldg(reinterpret_cast<const float4*>(p));
    if (MODE == 4) return hot ? ld_hint_na(p, pl) : ld_noalloc(p);
    if (MODE == 5) return hot ? ld_hint(p, pl) : ld_noalloc(p);
    if (MODE == 6) return ld_hint_na(p, hot ? pl : pf);   // pf = evict_normal in this mode
    return ld_hint(p, hot ? pl : pf);
}

struct ArcStream {
    const int32_t* __restrict__ colidx;
    int64_t limit, base;
    int32_t cur, nxt;
    __device__ __forceinline__ int32_t fetch(int64_t p) const {
        return p < limit ? __ldcs(colidx + p) : 0;
    }
    __device__ __forceinline__ void open(int64_t pos, int lane) {
        base = pos & ~int64_t(31);
        cur = fetch(base + lane);
        nxt = fetch(base + 32 + lane);
    }
    __device__ __forceinline__ void seek(int64_t pos, int lane) {
        if (pos < base + 32) return;
        if (pos < base + 64) { cur = nxt; base += 32; nxt = fetch(base + 32 + lane); }
        else open(pos, lane);
    }
};

// ---- LDG family ---------------------------------------------------------------------------
template <int LPR, int U, int MINB, int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS, MINB) ldg_kernel(const Args a) {
    constexpr int G = 32 / LPR;
    constexpr int D = LPR * 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LPR, grp = lane / LPR;
    const float* xcol = a.X + sub * 4;
    uint64_t pl = 0, pf = 0;
    if (MODE >= 2) { pl = policy_evict_last(); pf = MODE == 6 ? policy_evict_normal() : policy_evict_first(); }

    const int64_t first = ((int64_t)blockIdx.x * (THREADS / 32) + warp) * a.rows_per_warp;
    if (first >= a.n_rows) return;
    const int nrows = (int)min((int64_t)a.rows_per_warp, a.n_rows - first);
    const int64_t rp = lane <= nrows ? __ldg(a.rowptr + first + lane) : 0;
    ArcStream s;
    s.colidx = a.colidx;
    s.limit = __shfl_sync(kFull, rp, nrows);
    s.open(__shfl_sync(kFull, rp, 0), lane);

    for (int r = 0; r < nrows; ++r) {
        const int64_t beg = __shfl_sync(kFull, rp, r), end = __shfl_sync(kFull, rp, r + 1);
        const int64_t deg = end - beg;
        if (deg > kHub) continue;
        float4 acc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        int64_t k = beg;
        while (k < end) {
            s.seek(k, lane);
            const int off = (int)(k - s.base);
            const int cnt = (int)min(end - k, (int64_t)(32 - off));
            for (int t = 0; t < cnt; t += G * U) {
                float4 v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int kk = t + u * G + grp;
                    const int32_t idx = __shfl_sync(kFull, s.cur, (off + kk) & 31);
                    v[u] = load_row<MODE>(xcol + (int64_t)idx * a.ldx, kk < cnt,
                                          idx < a.hot_rows, pl, pf);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w;
                }
            }
            k += cnt;
        }
#pragma unroll
        for (int u = 1; u < U; ++u) {
            acc[0].x += acc[u].x; acc[0].y += acc[u].y; acc[0].z += acc[u].z; acc[0].w += acc[u].w;
        }
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
            acc[0].x += __shfl_xor_sync(kFull, acc[0].x, o);
            acc[0].y += __shfl_xor_sync(kFull, acc[0].y, o);
            acc[0].z += __shfl_xor_sync(kFull, acc[0].z, o);
            acc[0].w += __shfl_xor_sync(kFull, acc[0].w, o);
        }
        float* o = a.out + (first + r) * (2 * D) + sub * 4;
        if (grp == 0) __stcs(reinterpret_cast<float4*>(o), acc[0]);
        if (grp == (G >= 2 ? 1 : 0)) {
            const float inv = deg > 0 ? 1.f / (float)deg : 0.f;
            __stcs(reinterpret_cast<float4*>(o + D),
                   make_float4(acc[0].x * inv, acc[0].y * inv, acc[0].z * inv, acc[0].w * inv));
        }
    }
}

// ---- bulk-copy family: cp.async.bulk (TMA 1-D) row gathers into a per-warp smem ring ----------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                         uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Each warp: ring of STAGES chunks x 32 rows x (LPR*16) bytes.
template <int LPR, int STAGES, int WARPS, int LU>
__global__ void __launch_bounds__(WARPS * 32) bulk_kernel(const Args a) {
    constexpr int G = 32 / LPR;
    constexpr int D = LPR * 4;
    constexpr int ROWB = D * 4;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LPR, grp = lane / LPR;
    unsigned char* ring = smem + (size_t)warp * STAGES * 32 * ROWB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * STAGES * 32 * ROWB) +
                     warp * STAGES;
    const uint32_t ring_u32 = smem_u32(ring), bars_u32 = smem_u32(bars);

    if (lane == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(bars_u32 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const int64_t first = ((int64_t)blockIdx.x * WARPS + warp) * a.rows_per_warp;
    if (first >= a.n_rows) return;
    const int nrows = (int)min((int64_t)a.rows_per_warp, a.n_rows - first);
    const int64_t rp = lane <= nrows ? __ldg(a.rowptr + first + lane) : 0;
    const int64_t sbeg = __shfl_sync(kFull, rp, 0), send = __shfl_sync(kFull, rp, nrows);
    const int nchunks = (int)((send - sbeg + 31) >> 5);

    auto issue = [&](int c) {
        const int st = c % STAGES;
        const int64_t cstart = sbeg + 32 * (int64_t)c;
        const int cnt = (int)min((int64_t)32, send - cstart);
        const uint32_t bar = bars_u32 + 8 * st;
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)cnt * ROWB);
        __syncwarp();
        if (lane < cnt) {
            const int32_t idx = __ldcs(a.colidx + cstart + lane);
            bulk_g2s(ring_u32 + (uint32_t)(st * 32 + lane) * ROWB, a.X + (int64_t)idx * a.ldx,
                     ROWB, bar);
        }
    };

    for (int c = 0; c < STAGES && c < nchunks; ++c) issue(c);

    int r = 0;
    int64_t beg = __shfl_sync(kFull, rp, 0), end = __shfl_sync(kFull, rp, 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);

    auto flush = [&]() {
        float4 t = acc;
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
            t.x += __shfl_xor_sync(kFull, t.x, o);
            t.y += __shfl_xor_sync(kFull, t.y, o);
            t.z += __shfl_xor_sync(kFull, t.z, o);
            t.w += __shfl_xor_sync(kFull, t.w, o);
        }
        const int64_t deg = end - beg;
        float* o = a.out + (first + r) * (2 * D) + sub * 4;
        if (grp == 0) __stcs(reinterpret_cast<float4*>(o), t);
        if (grp == (G >= 2 ? 1 : 0)) {
            const float inv = deg > 0 ? 1.f / (float)deg : 0.f;
            __stcs(reinterpret_cast<float4*>(o + D),
                   make_float4(t.x * inv, t.y * inv, t.z * inv, t.w * inv));
        }
        acc = make_float4(0.f, 0.f, 0.f, 0.f);
        ++r;
        beg = end;
        end = __shfl_sync(kFull, rp, min(r + 1, nrows));
    };

    for (int c = 0; c < nchunks; ++c) {
        const int st = c % STAGES;
        mbar_wait(bars_u32 + 8 * st, (uint32_t)((c / STAGES) & 1));
        const int64_t cstart = sbeg + 32 * (int64_t)c;
        const int64_t cend = min(cstart + 32, send);
        const unsigned char* stage = ring + (size_t)st * 32 * ROWB + sub * 16;
        int64_t k = cstart;
        while (k < cend) {
            while (end <= k && r < nrows) flush();      // finished (or empty) rows
            const int64_t pend = min(end, cend);
            const int j0 = (int)(k - cstart), cnt = (int)(pend - k);
            for (int t = 0; t < cnt; t += G * LU) {
                float4 v[LU];
#pragma unroll
                for (int u = 0; u < LU; ++u) {
                    const int kk = t + u * G + grp;
                    v[u] = kk < cnt ? *reinterpret_cast<const float4*>(stage + (size_t)(j0 + kk) * ROWB)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < LU; ++u) {
                    acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
                }
            }
            k = pend;
        }
        __syncwarp();
        if (c + STAGES < nchunks) issue(c + STAGES);
    }
    while (r < nrows) flush();
}

template <typename K>
float time_kernel(K launch, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        printf("CUDA error: %s\n", cudaGetErrorString(err));
        return -1.f;
    }
    return ms / reps;
}

template <int LPR, int U, int MINB, int MODE, int THREADS>
float run_ldg(const Args& a, int reps) {
    const int64_t per_block = (int64_t)(THREADS / 32) * a.rows_per_warp;
    const unsigned grid = (unsigned)((a.n_rows + per_block - 1) / per_block);
    return time_kernel([&] { ldg_kernel<LPR, U, MINB, MODE, THREADS><<<grid, THREADS>>>(a); }, reps);
}

template <int LPR, int STAGES, int WARPS, int LU>
float run_bulk(const Args& a, int reps) {
    const size_t smem = (size_t)WARPS * STAGES * 32 * LPR * 16 + (size_t)WARPS * STAGES * 8;
    cudaFuncSetAttribute(bulk_kernel<LPR, STAGES, WARPS, LU>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t per_block = (int64_t)WARPS * a.rows_per_warp;
    const unsigned grid = (unsigned)((a.n_rows + per_block - 1) / per_block);
    return time_kernel(
        [&] { bulk_kernel<LPR, STAGES, WARPS, LU><<<grid, WARPS * 32, smem>>>(a); }, reps);
}

}  // namespace

// variant ids: family*1000 + ...  (see tools/exp_refex.py for the table)
extern "C" float exp_run(int variant, int d, const int64_t* rowptr, const int32_t* colidx,
                         const float* X, int64_t ldx, int64_t n_rows, float* out,
                         int rows_per_warp, int hot_rows, int reps) {
    Args a{rowptr, colidx, X, ldx, n_rows, out, rows_per_warp, hot_rows};
#define LDG(id, U, MINB, MODE, T)                                         \
    if (variant == id) return d == 64 ? run_ldg<16, U, MINB, MODE, T>(a, reps) \
                                      : run_ldg<8, U, MINB, MODE, T>(a, reps);
#define BULK(id, S, W, LU)                                                \
    if (variant == id) return d == 64 ? run_bulk<16, S, W, LU>(a, reps)    \
                                      : run_bulk<8, S, W, LU>(a, reps);
    LDG(0, 4, 1, 0, 256)    // library default at the time of writing
    LDG(1, 2, 1, 0, 256)
    LDG(2, 1, 1, 0, 256)
    LDG(3, 2, 4, 0, 256)
    LDG(4, 2, 5, 0, 256)
    LDG(5, 2, 6, 0, 256)
    LDG(6, 1, 6, 0, 256)
    LDG(7, 1, 8, 0, 256)
    LDG(8, 2, 8, 0, 256)
    LDG(9, 3, 4, 0, 256)
    LDG(10, 2, 8, 0, 128)
    LDG(11, 2, 10, 0, 128)
    LDG(12, 2, 12, 0, 128)
    LDG(13, 1, 16, 0, 128)
    LDG(20, 2, 4, 1, 256)   // L1::no_allocate
    LDG(21, 2, 5, 1, 256)
    LDG(22, 4, 1, 1, 256)
    LDG(30, 2, 4, 2, 256)   // hot rows evict_last, others default
    LDG(31, 2, 4, 3, 256)   // hot rows evict_last, others evict_first
    LDG(32, 2, 5, 2, 256)
    LDG(33, 2, 5, 3, 256)
    LDG(34, 2, 5, 4, 256)
    LDG(35, 2, 5, 5, 256)
    LDG(36, 1, 8, 4, 256)
    LDG(37, 3, 4, 4, 256)
    LDG(38, 2, 10, 4, 128)
    LDG(39, 1, 8, 6, 256)
    BULK(100, 2, 8, 4)
    BULK(101, 3, 8, 4)
    BULK(102, 2, 4, 4)
    BULK(103, 4, 4, 4)
    BULK(104, 2, 8, 8)
    BULK(105, 1, 8, 4)
    BULK(106, 2, 2, 4)
    BULK(107, 4, 2, 4)
    BULK(108, 8, 2, 4)
    printf("unknown variant %d\n", variant);
    return -2.f;
}
