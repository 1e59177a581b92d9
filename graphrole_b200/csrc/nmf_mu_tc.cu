// Path B, tensor-core path: one fused pass over X per NMF multiplicative-update iteration,
// hand-written for sm_100a with TMA (cp.async.bulk.tensor), mbarrier pipelines, tcgen05.mma
// (kind::tf32, fp32 accumulation in TMEM) and tcgen05.ld epilogues.
//
// Per 64-row block b of X (persistent CTAs, one per SM, blocks round-robin):
//   P1  XHt[64, 32]      = X_b (64 x f, K-major)  .  H^T      tcgen05.mma M=64  N=32 K=8 x f/8
//   E   W_b             *= XHt / (W_b (H H^T))               epilogue warps: tcgen05.ld, fp32 math,
//                                                            W_b -> global, tf32(W_b)^T -> smem
//   P2  (W^T X)^T[f, 32] += X_b^T (MN-major)     .  W_b      tcgen05.mma M=128 N=32 K=8 x 8 per
//                                                            128 columns; accumulators stay in TMEM
//       (W^T W)[32, 32]  += W_b^T                 .  W_b      tcgen05.mma M=64  N=32 K=8 x 8
// The accumulators of P2 live in TMEM for the whole kernel and are written once per CTA as
// partials; nmf_finish_iteration (nmf_mu.cu) reduces them in fixed order and updates H.
//
// X is read from HBM once per iteration: P2 re-loads the block's tiles through TMA a few
// microseconds after P1 touched them, i.e. from L2 (126 MB), with the 32-byte-atom 128B swizzle
// tcgen05 requires for MN-major tf32 operands; P1 uses the ordinary 128B swizzle (K-major).
// Algorithmic bytes per iteration: n*f*4 + 2*n*r*4.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue
// (TMEM lane quarter = warp_id % 4).  Software pipeline of the MMA warp:
//   P1(0); then for i >= 1: P1(i), P2(i-1); finally P2(last)     -- the epilogue of block i-1
// overlaps P1(i).  The shared-memory ring holds stages of 64 rows x 128 columns (4 TMA boxes).

#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "nmf_handle.cuh"

using namespace gr;

namespace {

constexpr int kThreads = 192;
constexpr int kBlockRows = 64;      // rows of X per block (UMMA M of P1)
constexpr int kRP = 32;             // roles padded (UMMA N)
constexpr int kBoxCols = 32;        // 32 fp32 = 128 B = one swizzle row
constexpr int kGroupCols = 128;     // columns per ring stage (UMMA M of P2)
constexpr int kBoxBytes = kBlockRows * kBoxCols * 4;        // 8 KB
constexpr int kStageBytes = 4 * kBoxBytes;                  // 32 KB
constexpr int kHBoxBytes = kRP * kBoxCols * 4;              // 4 KB
constexpr int kWnewBytes = kRP * kBlockRows * 4;            // 8 KB  (W_b^T, K-major: [role][row])
constexpr int kMaxGroups = 6;                               // f <= 768
constexpr float kEps = 1.1920928955078125e-07f;

// TMEM column map (512 columns allocated)
constexpr int kColD1 = 0;      // 2 x 32: XHt double buffer (M=64 layout)
constexpr int kColWtW = 64;    // 32: W^T W (M=64 layout, rows 0..31 valid)
constexpr int kColD2 = 96;     // groups x 32: (W^T X)^T, M=128 layout
constexpr int kTmemCols = 512;

struct TcParams {
    int64_t n;
    int f, r;
    int groups;          // ceil(f / 128)
    int stages;          // ring depth
    int64_t n_blocks;    // ceil(n / 64)
    const float* hht;    // [r, r]
    float* W;            // [n, r]
    float* part_wtx;     // [grid, 32, f]
    float* part_wtw;     // [grid, 32, r]
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// 32 lanes x 32 columns of fp32: thread t of the warp gets lane (base + t), columns c..c+31
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, float (&v)[32]) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]),
          "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]),
          "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]),
          "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]),
          "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// ---- descriptors (cute/arch/mma_sm100_desc.hpp bit layout) --------------------------------------
// shared-memory matrix descriptor: start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46),
// version = 1 [46,48), layout type [61,64)
constexpr uint64_t kLayoutSw128 = 2, kLayoutSw128Base32 = 1;
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo,
                                              uint64_t layout) {
    return (uint64_t)((addr & 0x3ffff) >> 4) | ((uint64_t)(lbo >> 4) << 16) |
           ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (layout << 61);
}
// instruction descriptor: D fp32 [4,6)=1, A/B tf32 [7,10)=[10,13)=2, a_major bit 15, b_major bit
// 16 (1 = MN-major), N >> 3 [17,23), M >> 4 [24,29)
constexpr uint32_t make_idesc(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
constexpr uint32_t kIdescP1 = make_idesc(64, kRP, 0, 0);    // X K-major, H K-major
constexpr uint32_t kIdescP2 = make_idesc(128, kRP, 1, 0);   // X^T MN-major, W_b^T K-major
constexpr uint32_t kIdescWtW = make_idesc(64, kRP, 0, 0);   // W_b^T K-major both sides

// ---- shared memory carve-up ---------------------------------------------------------------------
struct SmemLayout {
    uint32_t h;        // groups * 4 boxes of 4 KB
    uint32_t ring;     // stages * 32 KB
    uint32_t wnew;     // 2 x 8 KB + 4 KB readable pad (W^T W reads rows 32..63 of "A")
    uint32_t hht;      // 32 x 32 fp32
    uint32_t bars;     // mbarriers
    uint32_t tmem_ptr;
    uint32_t total;
};
__host__ __device__ inline SmemLayout smem_layout(int groups, int stages) {
    SmemLayout L;
    uint32_t off = 0;
    L.h = off;      off += (uint32_t)groups * 4 * kHBoxBytes;
    L.ring = off;   off += (uint32_t)stages * kStageBytes;
    L.wnew = off;   off += 2 * kWnewBytes + 4096;
    L.hht = off;    off += kRP * kRP * 4;
    L.bars = off;   off += 64 * 8;
    L.tmem_ptr = off; off += 16;
    L.total = off;
    return L;
}
// barrier slots
enum { B_FULL = 0, B_EMPTY = 8, B_HFULL = 16, B_D1FULL = 17, B_D1EMPTY = 19, B_WFULL = 21,
       B_WEMPTY = 23, B_D2FULL = 25, B_COUNT = 26 };

__global__ void __launch_bounds__(kThreads, 1)
nmf_fused_tc_kernel(const __grid_constant__ CUtensorMap map_x_k,
                    const __grid_constant__ CUtensorMap map_x_mn,
                    const __grid_constant__ CUtensorMap map_h, const TcParams p) {
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment for the 128B swizzle atoms
    unsigned char* smem = reinterpret_cast<unsigned char*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const SmemLayout L = smem_layout(p.groups, p.stages);
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_h = s_base + L.h, s_ring = s_base + L.ring, s_wnew = s_base + L.wnew;
    const uint32_t s_bars = s_base + L.bars;
    float* hht_s = reinterpret_cast<float*>(smem + L.hht);
    volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + L.tmem_ptr);
    auto bar = [&](int slot) { return s_bars + 8u * (uint32_t)slot; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = p.groups, S = p.stages;
    // blocks of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int64_t nb = (p.n_blocks - blockIdx.x + gridDim.x - 1) / gridDim.x;

    // ---- setup ----
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(bar(B_FULL + s), 1); mbar_init(bar(B_EMPTY + s), 1); }
        mbar_init(bar(B_HFULL), 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(B_D1FULL + i), 1);
            mbar_init(bar(B_D1EMPTY + i), 4);
            mbar_init(bar(B_WFULL + i), 4);
            mbar_init(bar(B_WEMPTY + i), 1);
        }
        mbar_init(bar(B_D2FULL), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < kRP * kRP; i += kThreads) {
        const int a = i / kRP, b = i % kRP;
        hht_s[i] = (a < p.r && b < p.r) ? p.hht[a * p.r + b] : 0.f;
    }
    // W^T tiles start as zeros (padded roles / rows stay zero), the pad must be finite
    for (int i = threadIdx.x; i < (2 * kWnewBytes + 4096) / 4; i += kThreads)
        reinterpret_cast<float*>(smem + L.wnew)[i] = 0.f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32((const void*)tmem_ptr_s)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // zero-fill visible to UMMA
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_s;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            mbar_expect_tx(bar(B_HFULL), (uint32_t)G * 4 * kHBoxBytes);
            for (int c = 0; c < G * 4; ++c)
                tma_load_2d(s_h + c * kHBoxBytes, &map_h, c * kBoxCols, 0, bar(B_HFULL));
            uint32_t it = 0;
            auto load_stage = [&](const CUtensorMap* map, int64_t blk, int g) {
                const int st = it % S;
                mbar_wait(bar(B_EMPTY + st), ((it / S) & 1) ^ 1);
                mbar_expect_tx(bar(B_FULL + st), kStageBytes);
                const int row = (int)(blk * kBlockRows);
                for (int c = 0; c < 4; ++c)
                    tma_load_2d(s_ring + st * kStageBytes + c * kBoxBytes, map,
                                g * kGroupCols + c * kBoxCols, row, bar(B_FULL + st));
                ++it;
            };
            for (int64_t i = 0; i <= nb; ++i) {
                if (i < nb)
                    for (int g = 0; g < G; ++g) load_stage(&map_x_k, blockIdx.x + i * gridDim.x, g);
                if (i >= 1)
                    for (int g = 0; g < G; ++g)
                        load_stage(&map_x_mn, blockIdx.x + (i - 1) * gridDim.x, g);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            mbar_wait(bar(B_HFULL), 0);
            uint32_t it = 0;
            for (int64_t i = 0; i <= nb; ++i) {
                if (i < nb) {
                    // ---- P1(i): XHt into D1[i & 1]
                    const int buf = (int)(i & 1);
                    mbar_wait(bar(B_D1EMPTY + buf), (uint32_t)(((i >> 1) & 1) ^ 1));
                    tc_fence_after();
                    const uint32_t d1 = tmem + kColD1 + buf * kRP;
                    for (int g = 0; g < G; ++g) {
                        const int st = it % S;
                        mbar_wait(bar(B_FULL + st), (it / S) & 1);
                        tc_fence_after();
                        for (int c = 0; c < 4; ++c) {
                            const uint32_t a0 = s_ring + st * kStageBytes + c * kBoxBytes;
                            const uint32_t b0 = s_h + (g * 4 + c) * kHBoxBytes;
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                tc_mma_tf32(d1, make_desc(a0 + k * 32, 16, 1024, kLayoutSw128),
                                            make_desc(b0 + k * 32, 16, 1024, kLayoutSw128),
                                            kIdescP1, (g | c | k) != 0);
                        }
                        tc_commit(bar(B_EMPTY + st));
                        ++it;
                    }
                    tc_commit(bar(B_D1FULL + buf));
                }
                if (i >= 1) {
                    // ---- P2(i-1): (W^T X)^T and W^T W with the updated W of block i-1
                    const int64_t j = i - 1;
                    const int buf = (int)(j & 1);
                    mbar_wait(bar(B_WFULL + buf), (uint32_t)((j >> 1) & 1));
                    tc_fence_after();
                    const uint32_t w0 = s_wnew + buf * kWnewBytes;
                    for (int g = 0; g < G; ++g) {
                        const int st = it % S;
                        mbar_wait(bar(B_FULL + st), (it / S) & 1);
                        tc_fence_after();
                        const uint32_t a0 = s_ring + st * kStageBytes;
#pragma unroll
                        for (int k = 0; k < 8; ++k)   // k-step = 8 rows of the block
                            tc_mma_tf32(tmem + kColD2 + g * kRP,
                                        make_desc(a0 + k * 1024, kBoxBytes, 512, kLayoutSw128Base32),
                                        make_desc(w0 + (k >> 2) * 4096 + (k & 3) * 32, 16, 1024,
                                                  kLayoutSw128),
                                        kIdescP2, (j | k) != 0);
                        tc_commit(bar(B_EMPTY + st));
                        ++it;
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint64_t d = make_desc(w0 + (k >> 2) * 4096 + (k & 3) * 32, 16, 1024,
                                                     kLayoutSw128);
                        tc_mma_tf32(tmem + kColWtW, d, d, kIdescWtW, (j | k) != 0);
                    }
                    tc_commit(bar(B_WEMPTY + buf));
                }
            }
            tc_commit(bar(B_D2FULL));
        }
    } else {
        // ================= epilogue warps (TMEM lane quarter q) =================
        const int q = warp & 3;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int r = p.r;
        for (int64_t i = 0; i < nb; ++i) {
            const int buf = (int)(i & 1);
            const uint32_t par = (uint32_t)((i >> 1) & 1);
            mbar_wait(bar(B_D1FULL + buf), par);
            tc_fence_after();
            float xht[32];
            tc_ld_32x32(tmem + lane_base + kColD1 + buf * kRP, xht);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_D1EMPTY + buf));

            // M=64 accumulator layout: row 16q + l lives in lane l < 16 of quarter q
            const int row_in_blk = q * 16 + lane;
            const int64_t row = (blockIdx.x + i * gridDim.x) * kBlockRows + row_in_blk;
            const bool valid = lane < 16 && row < p.n;
            float w[32], den[32];
#pragma unroll
            for (int l = 0; l < 32; ++l) {
                w[l] = (valid && l < r) ? __ldg(p.W + row * r + l) : 0.f;
                den[l] = 0.f;
            }
#pragma unroll
            for (int l = 0; l < 32; ++l) {
                const float wl = w[l];
#pragma unroll
                for (int jj = 0; jj < 32; jj += 4) {
                    const float4 hv = *reinterpret_cast<const float4*>(hht_s + l * kRP + jj);
                    den[jj] = fmaf(wl, hv.x, den[jj]);
                    den[jj + 1] = fmaf(wl, hv.y, den[jj + 1]);
                    den[jj + 2] = fmaf(wl, hv.z, den[jj + 2]);
                    den[jj + 3] = fmaf(wl, hv.w, den[jj + 3]);
                }
            }
            // the MMA of block i-2 must be done with this W^T buffer
            mbar_wait(bar(B_WEMPTY + buf), par ^ 1);
            unsigned char* wt = smem + L.wnew + buf * kWnewBytes;
#pragma unroll
            for (int l = 0; l < 32; ++l) {
                float d = den[l] == 0.f ? kEps : den[l];
                const float wn = (valid && l < r) ? w[l] * (xht[l] / d) : 0.f;
                if (valid && l < r) p.W[row * r + l] = wn;
                if (lane < 16) {
                    // K-major SW128 tile [role l][row k]: atom k/32, 128 B per role row,
                    // 16-byte chunk index XOR (l % 8)
                    const int k = row_in_blk;
                    const uint32_t off = (uint32_t)(k >> 5) * 4096 + (uint32_t)l * 128 +
                                         ((((uint32_t)(k & 31) >> 2) ^ ((uint32_t)l & 7)) << 4) +
                                         ((uint32_t)k & 3) * 4;
                    *reinterpret_cast<float*>(wt + off) = to_tf32(wn);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_WFULL + buf));
        }

        // ---- final: dump the TMEM accumulators as this CTA's partials
        mbar_wait(bar(B_D2FULL), 0);
        tc_fence_after();
        float v[32];
        for (int g = 0; g < G; ++g) {
            tc_ld_32x32(tmem + lane_base + kColD2 + g * kRP, v);
            const int col = g * kGroupCols + q * 32 + lane;   // M=128 layout: lane = row of D
            if (col < p.f)
#pragma unroll
                for (int l = 0; l < 32; ++l)
                    if (l < r)
                        p.part_wtx[((int64_t)blockIdx.x * kRP + l) * p.f + col] = v[l];
        }
        tc_ld_32x32(tmem + lane_base + kColWtW, v);
        if (q < 2 && lane < 16) {
            const int role = q * 16 + lane;   // M=64 layout
            if (role < r)
#pragma unroll
                for (int l = 0; l < 32; ++l)
                    if (l < r) p.part_wtw[((int64_t)blockIdx.x * kRP + role) * r + l] = v[l];
        }
        tc_fence_before();
    }

    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

// ---- host side ------------------------------------------------------------------------------------
struct TcState {
    int grid = 0;
    int stages = 0;
    size_t smem_bytes = 0;
    float* d_part_wtx = nullptr;
    float* d_part_wtw = nullptr;
    CUtensorMap map_x_k, map_x_mn, map_h;
    const float* X = nullptr;   // what the X maps were encoded for
    int64_t ldx = 0;
    const float* H = nullptr;
};

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) !=
                cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

int encode_2d(CUtensorMap* map, const float* base, uint64_t cols, uint64_t rows,
              uint64_t row_stride_bytes, uint32_t box_cols, uint32_t box_rows,
              CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(GR_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {row_stride_bytes};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t elem[2] = {1, 1};
    static const bool round_tf32 = [] {
        const char* s = getenv("GR_NMF_TMA_TF32");
        return !(s && s[0] == '0');
    }();
    const CUresult rc = fn(map, round_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32
                                           : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                           2, const_cast<float*>(base), dims, strides, box, elem,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return fail(GR_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
    return GR_OK;
}

}  // namespace

bool gr::nmf_tc_supported(const gr_nmf* h, const float* X, int64_t ldx) {
    if (getenv("GR_NMF_DISABLE_TC")) return false;
    return h->r <= kRP && h->f % 4 == 0 && ldx % 4 == 0 && aligned16(X) &&
           ceil_div(h->f, kGroupCols) <= kMaxGroups && h->n < ((int64_t)1 << 31) &&
           encode_fn() != nullptr;
}

int gr::nmf_iteration_tc(gr_nmf* h, const float* X, int64_t ldx, float* W, float* H,
                         cudaStream_t st) {
    TcState* s = static_cast<TcState*>(h->tc_state);
    const int groups = ceil_div(h->f, kGroupCols);
    if (!s) {
        s = new (std::nothrow) TcState();
        if (!s) return fail(GR_ERR_OUT_OF_MEMORY, "nmf tc state");
        h->tc_state = s;
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
        const int64_t n_blocks = ceil_div<int64_t>(h->n, kBlockRows);
        s->grid = (int)std::min<int64_t>(sms, n_blocks);
        int max_smem = 0;
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
        for (s->stages = 8; s->stages >= 2; --s->stages)
            if ((size_t)smem_layout(groups, s->stages).total + 1024 <= (size_t)max_smem) break;
        if (s->stages < 2) return fail(GR_ERR_CUDA, "nmf tc: shared memory budget too small");
        s->smem_bytes = (size_t)smem_layout(groups, s->stages).total + 1024;
        GR_CUDA_TRY(cudaFuncSetAttribute(nmf_fused_tc_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)s->smem_bytes));
        GR_CUDA_TRY(cudaMalloc(&s->d_part_wtx, (size_t)s->grid * kRP * h->f * sizeof(float)));
        GR_CUDA_TRY(cudaMalloc(&s->d_part_wtw, (size_t)s->grid * kRP * h->r * sizeof(float)));
    }
    if (s->X != X || s->ldx != ldx) {
        if (int rc = encode_2d(&s->map_x_k, X, (uint64_t)h->f, (uint64_t)h->n, (uint64_t)ldx * 4,
                               kBoxCols, kBlockRows, CU_TENSOR_MAP_SWIZZLE_128B))
            return rc;
        if (int rc = encode_2d(&s->map_x_mn, X, (uint64_t)h->f, (uint64_t)h->n, (uint64_t)ldx * 4,
                               kBoxCols, kBlockRows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
            return rc;
        s->X = X;
        s->ldx = ldx;
    }
    if (s->H != H) {
        if (int rc = encode_2d(&s->map_h, H, (uint64_t)h->f, (uint64_t)h->r, (uint64_t)h->f * 4,
                               kBoxCols, kRP, CU_TENSOR_MAP_SWIZZLE_128B))
            return rc;
        s->H = H;
    }

    if (int rc = nmf_hht(h, H, st)) return rc;
    TcParams p;
    p.n = h->n;
    p.f = h->f;
    p.r = h->r;
    p.groups = groups;
    p.stages = s->stages;
    p.n_blocks = ceil_div<int64_t>(h->n, kBlockRows);
    p.hht = h->d_hht;
    p.W = W;
    p.part_wtx = s->d_part_wtx;
    p.part_wtw = s->d_part_wtw;
    nmf_fused_tc_kernel<<<s->grid, kThreads, s->smem_bytes, st>>>(s->map_x_k, s->map_x_mn,
                                                                 s->map_h, p);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return fail(GR_ERR_CUDA, "nmf_fused_tc_kernel launch failed: %s", cudaGetErrorString(e));
    return nmf_finish_iteration(h, s->d_part_wtx, s->d_part_wtw, s->grid, kRP, H, st);
}

void gr::nmf_tc_release(gr_nmf* h) {
    TcState* s = static_cast<TcState*>(h->tc_state);
    if (!s) return;
    cudaFree(s->d_part_wtx);
    cudaFree(s->d_part_wtw);
    delete s;
    h->tc_state = nullptr;
}
