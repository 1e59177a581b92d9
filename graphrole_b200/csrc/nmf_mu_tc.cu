// Path B, tensor-core path: one fused pass over X per NMF multiplicative-update iteration,
// hand-written for sm_100a with TMA (cp.async.bulk.tensor), mbarrier pipelines, tcgen05.mma
// (kind::tf32, fp32 accumulation in TMEM) and tcgen05.ld epilogues.
//
// Per 64-row block b of X (persistent CTAs, one per SM, blocks round-robin):
//   P1  XHt[64, 32]      = X_b (64 x f, K-major)  .  H^T      tcgen05.mma M=64  N=32 K=8 x f/8
//   P1' Den[64, 32]      = W_b (64 x 32, K-major)  .  (H H^T)  tcgen05.mma M=64  N=32 K=8 x 4
//   E   W_b             *= XHt / Den                         epilogue warps: tcgen05.ld, 32 fp32
//                                                            ops per row; W_b tiles move by TMA
//                                                            (load + store), tf32(W_b) -> smem
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
//   * an epilogue that computes W (H H^T) per row on CUDA cores (1024 FMA + 32 IEEE divisions
//     per thread, one warp per scheduler) takes ~10 us per block and bounds the kernel
//     (timeline trace, GR_NMF_TRACE); the r x r product therefore also runs on the tensor
//     core (P1'), W tiles are moved by TMA, and the epilogue is ~250 instructions per row.

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
constexpr int kMaxTiles = 11;                               // f <= 704 (TMEM columns)
constexpr int kWSubBytes = 16 * kRP * 4;                    // 2 KB: one warp's 16 rows of a W tile
constexpr int kMaxStagesA = 5, kMaxStagesB = 4;
constexpr float kEps = 1.1920928955078125e-07f;

// TMEM column map (512 columns allocated)
constexpr int kColD1 = 0;      // 2 x 32: XHt double buffer (M=64 layout)
constexpr int kColDen = 64;    // 2 x 32: W (H H^T) double buffer (M=64 layout)
constexpr int kColWtW = 128;   // 32: W^T W (M=64 layout, rows 0..31 valid, 32..63 duplicates)
constexpr int kColD2 = 160;    // tiles x 32: (W^T X)^T, M=64 layout per 64-column tile
constexpr int kTmemCols = 512;

struct TcParams {
    int64_t n;
    int f, r;
    int groups;          // ceil(f / 128): P1 stages per block (H holds 4 * groups boxes)
    int tiles;           // ceil(f / 64): P2 stages per block
    int ring_a;          // P1 ring depth in 32 KB stages
    int ring_b;          // P2 ring depth in 16 KB stages
    int64_t n_blocks;    // ceil(n / 64)
    float* part_wtx;     // [grid, 32, f]
    float* part_wtw;     // [grid, 32, r]
    unsigned long long* trace;  // development: [4 roles][64 blocks][8 events] globaltimer ns (CTA 0)
    int debug;           // development switches (GR_NMF_TC_DEBUG): 1 skip P1 MMAs, 2 skip P2 MMAs,
                         // 4 skip epilogue math, 8 skip P2 TMA loads + MMAs, 16 experiment: P1 reads
                         // its K-major operand from tiles loaded with the 32-byte-atom swizzle
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
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {   // non-blocking
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int x, int y, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(map), "r"(x), "r"(y), "r"(src) : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
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
constexpr uint32_t kIdescP2 = make_idesc(64, kRP, 1, 1);    // X^T MN-major, W_b MN-major
constexpr uint32_t kIdescWtW = make_idesc(64, kRP, 1, 1);   // W_b MN-major on both sides

// ---- shared memory carve-up ---------------------------------------------------------------------
struct SmemLayout {
    uint32_t h;        // 4 * groups boxes of 4 KB
    uint32_t ring_a;   // P1 ring: stages of 32 KB
    uint32_t ring_b;   // P2 ring: stages of 16 KB
    uint32_t wio;      // 2 x 8 KB: W_b tile in (TMA load, fp32, K-major SW128) / out (TMA store)
    uint32_t wnew;     // 8 KB: tf32(W_b new), [row][role] in the 32-byte-atom 128B swizzle
    uint32_t hht;      // 4 KB: H H^T as a K-major SW128 tile (TMA)
    uint32_t bars;     // mbarriers
    uint32_t tmem_ptr;
    uint32_t total;
};
__host__ __device__ inline SmemLayout smem_layout(int groups, int ring_a, int ring_b) {
    SmemLayout L;
    uint32_t off = 0;
    L.h = off;      off += (uint32_t)groups * kStageABoxes * kHBoxBytes;
    off = (off + 1023u) & ~1023u;
    L.ring_a = off; off += (uint32_t)ring_a * kStageABytes;
    L.ring_b = off; off += (uint32_t)ring_b * kStageBytes;
    L.wio = off;    off += 2 * kWnewBytes;
    L.wnew = off;   off += kWnewBytes;
    L.hht = off;    off += kHBoxBytes;
    L.bars = off;   off += 64 * 8;
    L.tmem_ptr = off; off += 16;
    L.total = off;
    return L;
}
// barrier slots
enum { B_FULL_A = 0, B_EMPTY_A = 8, B_FULL_B = 16, B_EMPTY_B = 20, B_HFULL = 24, B_D1FULL = 25,
       B_D1EMPTY = 27, B_WFULL = 29, B_WEMPTY = 30, B_D2FULL = 31, B_WINFULL = 32, B_COUNT = 34 };

__global__ void __launch_bounds__(kThreads, 1)
nmf_fused_tc_kernel(const __grid_constant__ CUtensorMap map_x_k,
                    const __grid_constant__ CUtensorMap map_x_mn,
                    const __grid_constant__ CUtensorMap map_h,
                    const __grid_constant__ CUtensorMap map_hht,
                    const __grid_constant__ CUtensorMap map_w, const TcParams p) {
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment for the 128B swizzle atoms
    unsigned char* smem = reinterpret_cast<unsigned char*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const SmemLayout L = smem_layout(p.groups, p.ring_a, p.ring_b);
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_h = s_base + L.h, s_ra = s_base + L.ring_a, s_rb = s_base + L.ring_b;
    const uint32_t s_wnew = s_base + L.wnew, s_wio = s_base + L.wio, s_hht = s_base + L.hht;
    const uint32_t s_bars = s_base + L.bars;
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
        mbar_init(bar(B_WINFULL), 4);
        mbar_init(bar(B_WINFULL + 1), 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32((const void*)tmem_ptr_s)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_s;

    if (warp == 0) {
        // ================= TMA producer, P1 stream (HBM): H once, then 128-column stages =========
        if (lane == 0) {
            mbar_expect_tx(bar(B_HFULL), (uint32_t)(G * kStageABoxes + 1) * kHBoxBytes);
            for (int c = 0; c < G * kStageABoxes; ++c)
                tma_load_2d(s_h + c * kHBoxBytes, &map_h, c * kBoxCols, 0, bar(B_HFULL));
            tma_load_2d(s_hht, &map_hht, 0, 0, bar(B_HFULL));
            uint32_t it = 0;
            for (int64_t i = 0; i < nb; ++i) {
                const int row = (int)((blockIdx.x + i * gridDim.x) * kBlockRows);
                GR_TRACE(0, i, 0);
                for (int g = 0; g < G; ++g, ++it) {
                    const int st = it % NA;
                    mbar_wait(bar(B_EMPTY_A + st), ((it / NA) & 1) ^ 1);
                    if (g < 7) GR_TRACE(0, i, 1 + g);
                    mbar_expect_tx(bar(B_FULL_A + st), kStageABytes);
                    for (int c = 0; c < kStageABoxes; ++c)
                        tma_load_2d(s_ra + st * kStageABytes + c * kBoxBytes,
                                    (p.debug & 16) ? &map_x_mn : &map_x_k,
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
            // Issue order: P1(0); then P1(i), P2(i-1) for i >= 1; finally P2(last).  Measured
            // alternatives (C5, r = 32): interleaving P2(i-1) stages into P1(i) 9.6 ms, polling
            // both rings and consuming whichever stage landed 9.6 - 9.9 ms, this order 8.4 ms --
            // with 128 KB of ring space the slot turn-around bounds both streams, and blocking
            // mbarrier waits have less latency than a polling loop.
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
                                tc_mma_tf32(d1, make_desc(a0 + k * 32, 16, 1024,
                                                          (p.debug & 16) ? kLayoutSw128Base32
                                                                         : kLayoutSw128),
                                            make_desc(b0 + k * 32, 16, 1024, kLayoutSw128),
                                            kIdescP1, (g | c | k) != 0);
                        }
                        tc_commit(bar(B_EMPTY_A + st));
                    }
                    // ---- P1'(i): Den = W_b . (H H^T) from the TMA-loaded W tile.  (Its load was
                    // issued after the epilogue of block i-2, which needed P2(i-3): already done.)
                    mbar_wait(bar(B_WINFULL + buf1), (uint32_t)((i >> 1) & 1));
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma_tf32(tmem + kColDen + buf1 * kRP,
                                    make_desc(s_wio + buf1 * kWnewBytes + k * 32, 16, 1024, kLayoutSw128),
                                    make_desc(s_hht + k * 32, 16, 1024, kLayoutSw128), kIdescP1,
                                    k != 0);
                    tc_commit(bar(B_D1FULL + buf1));
                    GR_TRACE(2, i, 2);
                }
                if (do_p2) {
                    // ---- P2(i-1): (W^T X)^T and W^T W with the updated W of block i-1
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
                                        make_desc(s_wnew + k * 1024, 0, 512, kLayoutSw128Base32),
                                        kIdescP2, (j | k) != 0);
                        tc_commit(bar(B_EMPTY_B + st));
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        // LBO = 0: the second 32-role group of "A" aliases the first (rows 32..63 of
                        // the accumulator duplicate rows 0..31 and are never read)
                        const uint64_t d = make_desc(s_wnew + k * 1024, 0, 512, kLayoutSw128Base32);
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
        // this warp's 16 rows of the W tile: rows 16q..16q+15 = two 1 KB swizzle atoms
        auto w_sub = [&](int buf) { return s_wio + buf * kWnewBytes + q * kWSubBytes; };
        auto load_w_sub = [&](int64_t i) {   // elected lane: TMA load of block i's sub-tile
            if (i >= nb) return;
            const int buf = (int)(i & 1);
            const int row = (int)((blockIdx.x + i * gridDim.x) * kBlockRows) + q * 16;
            mbar_expect_tx(bar(B_WINFULL + buf), kWSubBytes);
            tma_load_2d(w_sub(buf), &map_w, 0, row, bar(B_WINFULL + buf));
        };
        if (lane == 0) { load_w_sub(0); load_w_sub(1); }
        const int k = q * 16 + (lane & 15);              // row of the block held by this lane
        for (int64_t i = 0; i < nb; ++i) {
            const int buf = (int)(i & 1);
            const uint32_t par = (uint32_t)((i >> 1) & 1);
            if (q == 0 && lane == 0) GR_TRACE(3, i, 0);
            // ---- XHt and Den from TMEM (M=64 layout: row 16q + l in lane l < 16 of quarter q)
            mbar_wait(bar(B_D1FULL + buf), par);
            mbar_wait(bar(B_WINFULL + buf), par);        // the W tile the MMA already consumed
            tc_fence_after();
            if (q == 0 && lane == 0) GR_TRACE(3, i, 1);
            float xht[32], den[32];
            tc_ld_32x32(tmem + lane_base + kColD1 + buf * kRP, xht);
            tc_ld_32x32(tmem + lane_base + kColDen + buf * kRP, den);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_D1EMPTY + buf));
            if (q == 0 && lane == 0) GR_TRACE(3, i, 2);
            // ---- W row from the tile: [row][role], 16-byte chunk c at c ^ (row % 8)
            unsigned char* wrow = smem + L.wio + buf * kWnewBytes + (uint32_t)k * 128;
            float wn[32];
            if (lane < 16) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 t = *reinterpret_cast<const float4*>(wrow + ((c ^ (k & 7)) << 4));
                    wn[4 * c] = t.x; wn[4 * c + 1] = t.y; wn[4 * c + 2] = t.z; wn[4 * c + 3] = t.w;
                }
#pragma unroll
                for (int l = 0; l < 32; ++l) {
                    const float d = den[l] == 0.f ? kEps : den[l];
                    wn[l] = l < r ? wn[l] * __fdividef(xht[l], d) : 0.f;
                }
                // fp32 result back into the tile (TMA store source)
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<float4*>(wrow + ((c ^ (k & 7)) << 4)) =
                        make_float4(wn[4 * c], wn[4 * c + 1], wn[4 * c + 2], wn[4 * c + 3]);
            }
            if (q == 0 && lane == 0) GR_TRACE(3, i, 3);
            // ---- tf32 copy for P2 / W^T W: [row][role], 32-byte chunk c at c ^ (row % 4).
            // P2(i-1) must be done with that tile.
            mbar_wait(bar(B_WEMPTY), (uint32_t)((i & 1) ^ 1));
            if (q == 0 && lane == 0) GR_TRACE(3, i, 4);
            if (lane < 16) {
                unsigned char* trow = smem + L.wnew + (uint32_t)k * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<float4*>(trow + ((((c >> 1) ^ (k & 3)) << 5) | ((c & 1) << 4))) =
                        make_float4(to_tf32(wn[4 * c]), to_tf32(wn[4 * c + 1]),
                                    to_tf32(wn[4 * c + 2]), to_tf32(wn[4 * c + 3]));
            }
            if (q == 0 && lane == 0) GR_TRACE(3, i, 5);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar(B_WFULL));
                // ---- W tile out (rows beyond n / roles beyond r are clipped by the tensor map),
                // then reuse the buffer for block i + 2
                const int row = (int)((blockIdx.x + i * gridDim.x) * kBlockRows) + q * 16;
                tma_store_2d(&map_w, 0, row, w_sub(buf));
                tma_store_commit_and_wait_read();
                load_w_sub(i + 2);
            }
            __syncwarp();
            if (q == 0 && lane == 0) GR_TRACE(3, i, 6);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");

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
    CUtensorMap map_x_k, map_x_mn, map_h, map_hht, map_w;
    const float* W = nullptr;
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
    return h->r <= kRP && h->r % 4 == 0 && h->f % 4 == 0 && ldx % 4 == 0 && aligned16(X) &&
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
        // P1 (HBM) ring: three 32 KB stages; P2 (L2) ring: two 16 KB stages; anything left goes
        // to the P2 ring first (its slots turn around faster)
        s->ring_a = 3;
        s->ring_b = 2;
        if (!fits(s->ring_a, s->ring_b)) s->ring_a = 2;
        if (!fits(s->ring_a, s->ring_b)) return fail(GR_ERR_CUDA, "nmf tc: shared memory budget");
        while (s->ring_b < kMaxStagesB && fits(s->ring_a, s->ring_b + 1)) ++s->ring_b;
        while (s->ring_a < kMaxStagesA && fits(s->ring_a + 1, s->ring_b)) ++s->ring_a;
        if (getenv("GR_NMF_RING_A") && getenv("GR_NMF_RING_B") &&
            fits(atoi(getenv("GR_NMF_RING_A")), atoi(getenv("GR_NMF_RING_B")))) {
            s->ring_a = std::max(1, std::min(kMaxStagesA, atoi(getenv("GR_NMF_RING_A"))));
            s->ring_b = std::max(1, std::min(kMaxStagesB, atoi(getenv("GR_NMF_RING_B"))));
        }
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
        if (int rc = encode_2d(&s->map_hht, h->d_hht, (uint64_t)h->r, (uint64_t)h->r,
                               (uint64_t)h->r * 4, kRP, kRP, CU_TENSOR_MAP_SWIZZLE_128B))
            return rc;
        s->H = H;
    }
    if (s->W != W) {
        // fp32 (no tf32 rounding: the tile is also the TMA store source), 16 rows per box
        if (int rc = encode_2d(&s->map_w, W, (uint64_t)h->r, (uint64_t)h->n, (uint64_t)h->r * 4,
                               kRP, 16, CU_TENSOR_MAP_SWIZZLE_128B, true))
            return rc;
        s->W = W;
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
    nmf_fused_tc_kernel<<<s->grid, kThreads, s->smem_bytes, st>>>(
        s->map_x_k, s->map_x_mn, s->map_h, s->map_hht, s->map_w, p);
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
