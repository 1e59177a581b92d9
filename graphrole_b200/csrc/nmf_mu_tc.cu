// Path B, tensor-core path: one fused pass over X per NMF multiplicative-update iteration,
// hand-written for sm_100a with TMA (cp.async.bulk.tensor), mbarrier pipelines, tcgen05.mma
// (kind::tf32, fp32 accumulation in TMEM) and tcgen05.ld epilogues.
//
// Per 64-row block b of X (persistent CTAs, one per SM, blocks round-robin):
//   P1  XHt[64, 32]      = X_b (64 x f, K-major)  .  H^T      tcgen05.mma M=64  N=32 K=8 x f/8
//   E   W_b             *= XHt / (W_b (H H^T))               epilogue warps: tcgen05.ld, fp32 math,
//                                                            W_b -> global, tf32(W_b)^T -> smem
//   P2  (W^T X)^T[f, 32] += X_b^T (MN-major)     .  W_b      tcgen05.mma M=64  N=32 K=8 x 8 per
//                                                            64 columns; accumulators stay in TMEM
//       (W^T W)[32, 32]  += W_b^T                 .  W_b      tcgen05.mma M=64  N=32 K=8 x 8
// The accumulators of P2 live in TMEM for the whole kernel and are written once per CTA as
// partials; nmf_finish_iteration (nmf_mu.cu) reduces them in fixed order and updates H.
//
// X is read from HBM once per iteration: P2 re-loads the block's tiles through TMA a few
// microseconds after P1 touched them, i.e. from L2 (126 MB), with the 32-byte-atom 128B swizzle
// tcgen05 requires for MN-major tf32 operands; P1 uses the ordinary 128B swizzle (K-major).
// Algorithmic bytes per iteration: n*f*4 + 2*n*r*4.
//
// Warp roles (224 threads): warp 0 = TMA producer of the P1 stream (HBM), warp 1 = MMA issuer,
// warps 2-5 = epilogue (TMEM lane quarter = warp_id % 4), warp 6 = TMA producer of the P2
// stream (L2 re-reads).  Each stream has its own shared-memory ring: P1 in 32 KB stages (four
// 64 x 32 boxes = 128 columns under one mbarrier), P2 in 16 KB stages (two boxes = the 64
// columns of one UMMA M tile).  MMA issue order: P1(0); then P1(i), P2(i-1) for i >= 1; finally
// P2(last) -- the epilogue of block i-1 overlaps P1(i) and the P1 ring keeps prefetching from
// HBM during P2(i-1).
//
// What was measured on B200 while getting here (profiles/README.md, C5 = 10 M x 512, r = 32):
//   * one in-order ring of 4 x 32 KB shared by both phases: 9.1 ms / iteration (both load
//     latencies exposed every block);
//   * per-box (8 KB) mbarrier stages: a single-thread producer/consumer handshake costs ~500
//     cycles, which caps a stream at ~4.4 TB/s regardless of ring depth; 32 KB per barrier
//     streams at 6.6-7.4 TB/s with only 3 stages (tools/exp_tma.cu);
//   * scalar per-row W loads/stores in the epilogue (16 rows x 4 B per warp instruction) cost
//     ~5 us per block; the W tile of a warp is contiguous in memory, so it is staged through
//     shared memory with coalesced 16-byte accesses instead.

#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "nmf_handle.cuh"

using namespace gr;

namespace {

constexpr int kThreads = 224;
constexpr int kBlockRows = 64;      // rows of X per block (UMMA M of P1)
constexpr int kRP = 32;             // roles padded (UMMA N)
constexpr int kBoxCols = 32;        // 32 fp32 = 128 B = one swizzle row
constexpr int kTileCols = 64;       // columns per P2 stage (UMMA M of P2)
constexpr int kBoxBytes = kBlockRows * kBoxCols * 4;        // 8 KB
constexpr int kStageABoxes = 4;                             // P1 stage: 128 columns
constexpr int kStageABytes = kStageABoxes * kBoxBytes;      // 32 KB
constexpr int kStageBytes = 2 * kBoxBytes;                  // 16 KB (P2 stage)
constexpr int kHBoxBytes = kRP * kBoxCols * 4;              // 4 KB
constexpr int kWnewBytes = kRP * kBlockRows * 4;            // 8 KB  (W_b^T, K-major: [role][row])
constexpr int kMaxTiles = 12;                               // f <= 768
constexpr int kMaxStagesA = 5, kMaxStagesB = 4;
constexpr int kEpiStageFloats = 16 * 33;                    // per-warp W tile staging (padded)
constexpr float kEps = 1.1920928955078125e-07f;

// TMEM column map (512 columns allocated)
constexpr int kColD1 = 0;      // 2 x 32: XHt double buffer (M=64 layout)
constexpr int kColWtW = 64;    // 32: W^T W (M=64 layout, rows 0..31 valid)
constexpr int kColD2 = 96;     // tiles x 32: (W^T X)^T, M=64 layout per 64-column tile
constexpr int kTmemCols = 512;

struct TcParams {
    int64_t n;
    int f, r;
    int groups;          // ceil(f / 128): P1 stages per block (H holds 4 * groups boxes)
    int tiles;           // ceil(f / 64): P2 stages per block
    int ring_a;          // P1 ring depth in 32 KB stages
    int ring_b;          // P2 ring depth in 16 KB stages
    int64_t n_blocks;    // ceil(n / 64)
    const float* hht;    // [r, r]
    float* W;            // [n, r]
    float* part_wtx;     // [grid, 32, f]
    float* part_wtw;     // [grid, 32, r]
    unsigned long long* trace;  // development: [4 roles][64 blocks][8 events] globaltimer ns (CTA 0)
    int debug;           // development switches (GR_NMF_TC_DEBUG): 1 skip P1 MMAs, 2 skip P2 MMAs,
                         // 4 skip epilogue math, 8 skip P2 TMA loads + MMAs
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
// L2 prefetch of a whole tensor box (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int x, int y) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];"
                 ::"l"(map), "r"(x), "r"(y) : "memory");
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
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define GR_TRACE(role, blk, ev)                                                             \
    do {                                                                                    \
        if (p.trace && blockIdx.x == 0 && (blk) < 64)                                       \
            p.trace[((role) * 64 + (blk)) * 8 + (ev)] = gtime();                            \
    } while (0)
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
constexpr uint32_t kIdescP2 = make_idesc(64, kRP, 1, 0);    // X^T MN-major, W_b^T K-major
constexpr uint32_t kIdescWtW = make_idesc(64, kRP, 0, 0);   // W_b^T K-major both sides

// ---- shared memory carve-up ---------------------------------------------------------------------
struct SmemLayout {
    uint32_t h;        // 4 * groups boxes of 4 KB
    uint32_t ring_a;   // P1 ring: stages of 32 KB
    uint32_t ring_b;   // P2 ring: stages of 16 KB
    uint32_t wnew;     // 8 KB W_b^T tile + 4 KB readable pad (W^T W reads rows 32..63 of "A")
    uint32_t hht;      // 32 x 32 fp32
    uint32_t epi;      // 4 warps x 16 x 33 floats
    uint32_t bars;     // mbarriers
    uint32_t tmem_ptr;
    uint32_t total;
};
__host__ __device__ inline SmemLayout smem_layout(int groups, int ring_a, int ring_b) {
    SmemLayout L;
    uint32_t off = 0;
    L.h = off;      off += (uint32_t)groups * 4 * kHBoxBytes;
    off = (off + 1023u) & ~1023u;
    L.ring_a = off; off += (uint32_t)ring_a * kStageABytes;
    L.ring_b = off; off += (uint32_t)ring_b * kStageBytes;
    L.wnew = off;   off += kWnewBytes + 4096;
    L.hht = off;    off += kRP * kRP * 4;
    L.epi = off;    off += 4 * kEpiStageFloats * 4;
    L.bars = off;   off += 64 * 8;
    L.tmem_ptr = off; off += 16;
    L.total = off;
    return L;
}
// barrier slots
enum { B_FULL_A = 0, B_EMPTY_A = 8, B_FULL_B = 16, B_EMPTY_B = 20, B_HFULL = 24, B_D1FULL = 25,
       B_D1EMPTY = 27, B_WFULL = 29, B_WEMPTY = 30, B_D2FULL = 31, B_COUNT = 32 };

__global__ void __launch_bounds__(kThreads, 1)
nmf_fused_tc_kernel(const __grid_constant__ CUtensorMap map_x_k,
                    const __grid_constant__ CUtensorMap map_x_mn,
                    const __grid_constant__ CUtensorMap map_h, const TcParams p) {
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment for the 128B swizzle atoms
    unsigned char* smem = reinterpret_cast<unsigned char*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const SmemLayout L = smem_layout(p.groups, p.ring_a, p.ring_b);
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_h = s_base + L.h, s_ra = s_base + L.ring_a, s_rb = s_base + L.ring_b;
    const uint32_t s_wnew = s_base + L.wnew;
    const uint32_t s_bars = s_base + L.bars;
    float* hht_s = reinterpret_cast<float*>(smem + L.hht);
    volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + L.tmem_ptr);
    auto bar = [&](int slot) { return s_bars + 8u * (uint32_t)slot; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = p.groups, T = p.tiles, NA = p.ring_a, NB = p.ring_b;
    // blocks of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int64_t nb = (p.n_blocks - blockIdx.x + gridDim.x - 1) / gridDim.x;

    // ---- setup ----
    if (threadIdx.x == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(bar(B_FULL_A + s), 1); mbar_init(bar(B_EMPTY_A + s), 1); }
        for (int s = 0; s < NB; ++s) { mbar_init(bar(B_FULL_B + s), 1); mbar_init(bar(B_EMPTY_B + s), 1); }
        mbar_init(bar(B_HFULL), 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(B_D1FULL + i), 1);
            mbar_init(bar(B_D1EMPTY + i), 4);
        }
        mbar_init(bar(B_WFULL), 4);
        mbar_init(bar(B_WEMPTY), 1);
        mbar_init(bar(B_D2FULL), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < kRP * kRP; i += kThreads) {
        const int a = i / kRP, b = i % kRP;
        hht_s[i] = (a < p.r && b < p.r) ? p.hht[a * p.r + b] : 0.f;
    }
    // the W^T tile starts as zeros (padded roles stay zero), the pad behind it must be finite
    for (int i = threadIdx.x; i < (kWnewBytes + 4096) / 4; i += kThreads)
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
        // ================= TMA producer, P1 stream (HBM): H once, then 128-column stages =========
        if (lane == 0) {
            mbar_expect_tx(bar(B_HFULL), (uint32_t)G * 4 * kHBoxBytes);
            for (int c = 0; c < G * 4; ++c)
                tma_load_2d(s_h + c * kHBoxBytes, &map_h, c * kBoxCols, 0, bar(B_HFULL));
            uint32_t it = 0;
            for (int64_t i = 0; i < nb; ++i) {
                const int row = (int)((blockIdx.x + i * gridDim.x) * kBlockRows);
                GR_TRACE(0, i, 0);
                for (int g = 0; g < G; ++g, ++it) {
                    const int st = it % NA;
                    mbar_wait(bar(B_EMPTY_A + st), ((it / NA) & 1) ^ 1);
                    if (g < 4) GR_TRACE(0, i, 1 + g);
                    mbar_expect_tx(bar(B_FULL_A + st), kStageABytes);
                    for (int c = 0; c < kStageABoxes; ++c)
                        tma_load_2d(s_ra + st * kStageABytes + c * kBoxBytes, &map_x_k,
                                    (g * kStageABoxes + c) * kBoxCols, row, bar(B_FULL_A + st));
                }
            }
        }
    } else if (warp == 6) {
        // ================= TMA producer, P2 stream (L2 re-reads, 32-byte-atom swizzle) ==========
        if (lane == 0 && !(p.debug & 8)) {
            uint32_t it = 0;
            for (int64_t i = 0; i < nb; ++i) {
                const int row = (int)((blockIdx.x + i * gridDim.x) * kBlockRows);
                for (int t = 0; t < T; ++t, ++it) {
                    const int st = it % NB;
                    mbar_wait(bar(B_EMPTY_B + st), ((it / NB) & 1) ^ 1);
                    if (t < 8) GR_TRACE(1, i, t);
                    mbar_expect_tx(bar(B_FULL_B + st), kStageBytes);
                    for (int c = 0; c < 2; ++c)
                        tma_load_2d(s_rb + st * kStageBytes + c * kBoxBytes, &map_x_mn,
                                    t * kTileCols + c * kBoxCols, row, bar(B_FULL_B + st));
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            mbar_wait(bar(B_HFULL), 0);
            uint32_t ita = 0, itb = 0;
            for (int64_t i = 0; i <= nb; ++i) {
                const bool do_p1 = i < nb, do_p2 = i >= 1;
                const int buf1 = (int)(i & 1);
                const int64_t j = i - 1;
                const uint32_t d1 = tmem + kColD1 + buf1 * kRP;
                if (do_p1) {
                    // ---- P1(i): X_b . H^T into D1, 128 columns (16 k-steps) per stage
                    GR_TRACE(2, i, 0);
                    mbar_wait(bar(B_D1EMPTY + buf1), (uint32_t)(((i >> 1) & 1) ^ 1));
                    tc_fence_after();
                    for (int g = 0; g < G; ++g, ++ita) {
                        const int st = ita % NA;
                        mbar_wait(bar(B_FULL_A + st), (ita / NA) & 1);
                        if (g == 0) GR_TRACE(2, i, 1);
                        tc_fence_after();
#pragma unroll
                        for (int c = 0; c < kStageABoxes; ++c) {
                            const uint32_t a0 = s_ra + st * kStageABytes + c * kBoxBytes;
                            const uint32_t b0 = s_h + (g * kStageABoxes + c) * kHBoxBytes;
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (!(p.debug & 1))
                                tc_mma_tf32(d1, make_desc(a0 + k * 32, 16, 1024, kLayoutSw128),
                                            make_desc(b0 + k * 32, 16, 1024, kLayoutSw128),
                                            kIdescP1, (g | c | k) != 0);
                        }
                        tc_commit(bar(B_EMPTY_A + st));
                    }
                    tc_commit(bar(B_D1FULL + buf1));
                    GR_TRACE(2, i, 2);
                }
                if (do_p2) {
                    // ---- P2(i-1): (W^T X)^T and W^T W with the updated W of block i-1.  Issued
                    // after P1(i), so the epilogue of block i-1 had all of P1(i) to finish.
                    mbar_wait(bar(B_WFULL), (uint32_t)(j & 1));
                    GR_TRACE(2, j, 3);
                    tc_fence_after();
                    for (int t = 0; t < T; ++t, ++itb) {
                        const int st = itb % NB;
                        if (!(p.debug & 8)) mbar_wait(bar(B_FULL_B + st), (itb / NB) & 1);
                        tc_fence_after();
                        const uint32_t a0 = s_rb + st * kStageBytes;
#pragma unroll
                        for (int k = 0; k < 8; ++k)   // k-step = 8 rows of the block
                            if (!(p.debug & 10))
                            tc_mma_tf32(tmem + kColD2 + t * kRP,
                                        make_desc(a0 + k * 1024, kBoxBytes, 512, kLayoutSw128Base32),
                                        make_desc(s_wnew + (k >> 2) * 4096 + (k & 3) * 32, 16, 1024,
                                                  kLayoutSw128),
                                        kIdescP2, (j | k) != 0);
                        tc_commit(bar(B_EMPTY_B + st));
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint64_t d = make_desc(s_wnew + (k >> 2) * 4096 + (k & 3) * 32, 16,
                                                     1024, kLayoutSw128);
                        tc_mma_tf32(tmem + kColWtW, d, d, kIdescWtW, (j | k) != 0);
                    }
                    tc_commit(bar(B_WEMPTY));
                    GR_TRACE(2, j, 4);
                }
            }
            tc_commit(bar(B_D2FULL));
        }
    } else {
        // ================= epilogue warps (TMEM lane quarter q) =================
        const int q = warp & 3;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int r = p.r;
        float* stg = reinterpret_cast<float*>(smem + L.epi) + q * kEpiStageFloats;   // [16][33]
        unsigned char* wt = smem + L.wnew;
        for (int64_t i = 0; i < nb; ++i) {
            const int buf = (int)(i & 1);
            // M=64 accumulator layout: row 16q + l of the block lives in lane l < 16 of quarter q
            const int64_t row0 = (blockIdx.x + i * gridDim.x) * kBlockRows + q * 16;
            const int rows_valid = (int)max((int64_t)0, min((int64_t)16, p.n - row0));
            const int n_el = rows_valid * r;                // floats of this warp's W tile
            float* wtile = p.W + row0 * r;                  // contiguous [rows_valid, r]
            if (q == 0 && lane == 0) GR_TRACE(3, i, 0);
            // ---- W tile -> staging (coalesced 16-byte loads; independent of the MMA)
            for (int e = lane * 4; e < 16 * r; e += 128) {
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (e + 3 < n_el) {
                    const float4 t = __ldcs(reinterpret_cast<const float4*>(wtile + e));
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                } else {
                    for (int u = 0; u < 4; ++u) if (e + u < n_el) v[u] = __ldcs(wtile + e + u);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) stg[((e + u) / r) * 33 + (e + u) % r] = v[u];
            }
            __syncwarp();
            float w[32], den[32];
#pragma unroll
            for (int l = 0; l < 32; ++l) {
                w[l] = (lane < 16 && l < r) ? stg[lane * 33 + l] : 0.f;
                den[l] = 0.f;
            }
            if (!(p.debug & 4))
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
            // ---- XHt from TMEM
            if (q == 0 && lane == 0) GR_TRACE(3, i, 1);
            mbar_wait(bar(B_D1FULL + buf), (uint32_t)((i >> 1) & 1));
            if (q == 0 && lane == 0) GR_TRACE(3, i, 2);
            tc_fence_after();
            float xht[32];
            tc_ld_32x32(tmem + lane_base + kColD1 + buf * kRP, xht);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_D1EMPTY + buf));
            // ---- W_new: to staging (for the coalesced store) and, tf32-rounded and transposed,
            // into the K-major SW128 tile P2 reads.  P2(i-1) must be done with that tile.
            if (q == 0 && lane == 0) GR_TRACE(3, i, 3);
            mbar_wait(bar(B_WEMPTY), (uint32_t)((i & 1) ^ 1));
            if (q == 0 && lane == 0) GR_TRACE(3, i, 4);
            const int k = q * 16 + lane;                    // row of the block = K index of P2
#pragma unroll
            for (int l = 0; l < 32; ++l) {
                const float d = den[l] == 0.f ? kEps : den[l];
                const float wn = (lane < rows_valid && l < r) ? w[l] * (xht[l] / d) : 0.f;
                if (lane < 16) {
                    if (l < r) stg[lane * 33 + l] = wn;
                    // tile [role l][row k]: atom k/32, 128 B per role row, 16-byte chunk ^ (l % 8)
                    const uint32_t off = (uint32_t)(k >> 5) * 4096 + (uint32_t)l * 128 +
                                         ((((uint32_t)(k & 31) >> 2) ^ ((uint32_t)l & 7)) << 4) +
                                         ((uint32_t)k & 3) * 4;
                    *reinterpret_cast<float*>(wt + off) = to_tf32(wn);
                }
            }
            if (q == 0 && lane == 0) GR_TRACE(3, i, 5);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_WFULL));
            if (q == 0 && lane == 0) GR_TRACE(3, i, 6);
            // ---- staging -> global (coalesced 16-byte stores)
            for (int e = lane * 4; e < 16 * r; e += 128) {
                float v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = stg[((e + u) / r) * 33 + (e + u) % r];
                if (e + 3 < n_el) {
                    __stcs(reinterpret_cast<float4*>(wtile + e), make_float4(v[0], v[1], v[2], v[3]));
                } else {
                    for (int u = 0; u < 4; ++u) if (e + u < n_el) __stcs(wtile + e + u, v[u]);
                }
            }
            __syncwarp();
            if (q == 0 && lane == 0) GR_TRACE(3, i, 7);
        }

        // ---- final: dump the TMEM accumulators as this CTA's partials
        mbar_wait(bar(B_D2FULL), 0);
        tc_fence_after();
        float v[32];
        for (int t = 0; t < T; ++t) {
            tc_ld_32x32(tmem + lane_base + kColD2 + t * kRP, v);
            const int col = t * kTileCols + q * 16 + lane;   // M=64 layout: row 16q + l in lane l
            if (lane < 16 && col < p.f)
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
    int ring_a = 0, ring_b = 0;
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
              CUtensorMapSwizzle swizzle, bool plain_f32 = false) {
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
    const CUresult rc = fn(map, (round_tf32 && !plain_f32) ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32
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
           ceil_div(h->f, kTileCols) <= kMaxTiles && h->n < ((int64_t)1 << 31) &&
           encode_fn() != nullptr;
}

int gr::nmf_iteration_tc(gr_nmf* h, const float* X, int64_t ldx, float* W, float* H,
                         cudaStream_t st) {
    TcState* s = static_cast<TcState*>(h->tc_state);
    const int groups = ceil_div(h->f, kStageABoxes * kBoxCols), tiles = ceil_div(h->f, kTileCols);
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
        // shared memory left after H goes to the rings: the P1 (HBM) ring first -- three 32 KB
        // stages stream at > 6.5 TB/s (tools/exp_tma.cu) -- then the P2 (L2) ring
        auto fits = [&](int a, int b) {
            return (size_t)smem_layout(groups, a, b).total + 1024 <= (size_t)max_smem;
        };
        s->ring_a = 3;
        s->ring_b = 2;
        if (!fits(s->ring_a, s->ring_b)) s->ring_a = 2;
        if (!fits(s->ring_a, s->ring_b)) return fail(GR_ERR_CUDA, "nmf tc: shared memory budget");
        while (s->ring_b < kMaxStagesB && fits(s->ring_a, s->ring_b + 1)) ++s->ring_b;
        while (s->ring_a < kMaxStagesA && fits(s->ring_a + 1, s->ring_b)) ++s->ring_a;
        if (const char* e = getenv("GR_NMF_RING_A")) s->ring_a = std::max(1, std::min(s->ring_a, atoi(e)));
        if (const char* e = getenv("GR_NMF_RING_B")) s->ring_b = std::max(1, std::min(s->ring_b, atoi(e)));
        s->smem_bytes = (size_t)smem_layout(groups, s->ring_a, s->ring_b).total + 1024;
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
    p.tiles = tiles;
    p.ring_a = s->ring_a;
    p.ring_b = s->ring_b;
    p.n_blocks = ceil_div<int64_t>(h->n, kBlockRows);
    p.hht = h->d_hht;
    p.W = W;
    p.part_wtx = s->d_part_wtx;
    p.part_wtw = s->d_part_wtw;
    p.trace = nullptr;
    static unsigned long long* d_trace = nullptr;
    if (getenv("GR_NMF_TRACE")) {
        if (!d_trace) cudaMalloc(&d_trace, 4 * 64 * 8 * sizeof(unsigned long long));
        cudaMemsetAsync(d_trace, 0, 4 * 64 * 8 * sizeof(unsigned long long), st);
        p.trace = d_trace;
    }
    p.debug = getenv("GR_NMF_TC_DEBUG") ? atoi(getenv("GR_NMF_TC_DEBUG")) : 0;
    nmf_fused_tc_kernel<<<s->grid, kThreads, s->smem_bytes, st>>>(s->map_x_k, s->map_x_mn,
                                                                 s->map_h, p);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return fail(GR_ERR_CUDA, "nmf_fused_tc_kernel launch failed: %s", cudaGetErrorString(e));
    if (p.trace) {
        static unsigned long long h_trace[4 * 64 * 8];
        cudaStreamSynchronize(st);
        cudaMemcpy(h_trace, d_trace, sizeof(h_trace), cudaMemcpyDeviceToHost);
        unsigned long long t0 = ~0ull;
        for (auto v : h_trace) if (v && v < t0) t0 = v;
        const char* names[4] = {"prodA", "prodB", "mma", "epi0"};
        const int first = getenv("GR_NMF_TRACE_FIRST") ? atoi(getenv("GR_NMF_TRACE_FIRST")) : 8;
        for (int b = first; b < first + 6; ++b)
            for (int role = 0; role < 4; ++role) {
                printf("blk %2d %-5s", b, names[role]);
                for (int ev = 0; ev < 8; ++ev) {
                    const unsigned long long v = h_trace[(role * 64 + b) * 8 + ev];
                    if (v) printf(" %8.2f", (double)(v - t0) / 1000.0); else printf("        -");
                }
                printf("\n");
            }
        fflush(stdout);
    }
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
