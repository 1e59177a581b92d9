// Path B, tensor-core path: one fused pass over X per NMF multiplicative-update iteration,
// hand-written for sm_100a with TMA (cp.async.bulk.tensor), mbarrier pipelines, tcgen05.mma
// (kind::tf32, fp32 accumulation in TMEM), tcgen05.ld epilogues and -- by default -- thread-block
// clusters of two CTAs that exchange partial products through distributed shared memory.
//
// Per 64-row block b of X (persistent CTAs, one per SM, blocks round-robin):
//   P1  XHt[64, 32]      = X_b (64 x f, K-major)  .  H^T      tcgen05.mma M=64  N=32 K=8 x f/8
//   P1' Den[64, 32]      = W_b (64 x 32, K-major)  .  (H H^T)  tcgen05.mma M=64  N=32 K=8 x 4
//   E   W_b             *= XHt / Den                         epilogue warps: tcgen05.ld.16x256b (all
//                                                            32 lanes hold data), 16 elements per
//                                                            thread; W_b tiles move by TMA (load +
//                                                            store), tf32(W_b) -> smem
//   P2  (W^T X)^T[f, 32] += X_b^T (MN-major)     .  W_b      tcgen05.mma M=128 N=32 K=8 x 8 per
//                                                            128 columns; accumulators stay in TMEM
//       (W^T W)[32, 32]  += W_b^T                 .  W_b      tcgen05.mma M=64  N=32 K=8 x 8
// The accumulators of P2 live in TMEM for the whole kernel and are written once per CTA as
// partials; nmf_finish_iteration (nmf_mu.cu) reduces them in fixed order and updates H.
//
// X is read from HBM once per iteration: P2 re-loads the block's tiles through TMA a few
// microseconds after P1 touched them, i.e. from L2 (126 MB), with the 32-byte-atom 128B swizzle
// tcgen05 requires for MN-major tf32 operands; P1 uses the ordinary 128B swizzle (K-major; the
// 32-byte-atom layout is refused for K-major operands, so one shared-memory copy cannot serve
// both).  Algorithmic bytes per iteration: n*f*4 + 2*n*r*4.
//
// Warp roles (224 threads): warp 0 = TMA producer of the P1 stream (HBM), warp 1 = MMA issuer,
// warps 2-5 = epilogue (TMEM lane quarter = warp_id % 4), warp 6 = TMA producer of the P2
// stream (L2 re-reads).  Each stream has its own shared-memory ring of 32 KB stages (four
// 64 x 32 boxes = 128 columns under one mbarrier).  MMA issue order: P1(i) then P2(i - LAG),
// LAG = 1 with one CTA per block, 2 with CTA pairs -- the epilogue of the earlier block overlaps
// P1(i) and the P1 ring keeps prefetching from HBM during P2.
//
// Column-split CTA pairs (CL = 2, the default when f > 128): two CTAs of a thread-block cluster
// work on the same 64-row block, each on half of the columns.  Every CTA then keeps only half of
// H (32 KB instead of 64 KB at f = 512) and a block's X tiles are 64 KB per CTA, so each ring
// holds a whole half-block ahead (with one CTA per block the fourth P1 stage of every block is
// loaded on demand, ~2 us of HBM latency exposed per block -- timeline traces in profiles/).
// The price: X H^T is a sum over columns, so each CTA sends its partial [64, 32] tile to the
// peer through distributed shared memory: st.async stores whose bytes complete an mbarrier of
// the peer (one hop, no fence), software-pipelined one block deep so the ~1.5 us hop is off the
// critical path.  Both CTAs do the (identical, fp32 addition is commutative) W update; rank 0
// stores W, the ranks alternate on W^T W, each rank stores its own columns of W^T X.
//
// What was measured on B200 while getting here (profiles/README.md, C5 = 10 M x 512, r = 32):
//   * one in-order ring of 4 x 32 KB shared by both phases: 9.1 ms / iteration (both load
//     latencies exposed every block);
//   * per-box (8 KB) mbarrier stages: a single-thread producer/consumer handshake costs ~500
//     cycles, which caps a stream at ~4.4 TB/s regardless of ring depth; 32 KB per barrier
//     streams at 6.6-7.4 TB/s with only 3 stages (tools/exp_tma.cu);
//   * an epilogue that computes W (H H^T) per row on CUDA cores takes ~10 us per block; the
//     r x r product therefore also runs on the tensor core (P1') and W tiles move by TMA;
//   * tcgen05.mma issued under `if (lane == 0)` compiles to an ELECT / BRA loop plus 17
//     instructions of descriptor arithmetic, ~105 cycles per MMA: the kernel was issue-bound at
//     8.45 ms; elect.sync + per-stage descriptors with constant increments -> 6.3 ms;
//   * an M=64 N=32 K=8 TF32 MMA costs ~50 cycles of tensor pipe, M=128 ~76: 128-column P2 tiles;
//   * 64 remote mbarrier arrives per block serialise (~1.5 us), a cluster-scope fence in front of
//     one arrive per warp is worse (10.3 ms); interleaving P2's accumulators k-outer doubles
//     P2's time; final: 4.3 - 4.9 ms per iteration (r = 4 .. 32), tensor-pipe bound.

#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "nmf_handle.cuh"

using namespace gr;

namespace {

constexpr int kThreads = 224;
constexpr int kBlockRows = 64;      // rows of X per block (UMMA M of P1)
constexpr int kRP = 32;             // roles padded (UMMA N)
constexpr int kBoxCols = 32;        // 32 fp32 = 128 B = one swizzle row
#ifndef GR_NMF_P2M
#define GR_NMF_P2M 128
#endif
constexpr int kP2M = GR_NMF_P2M;    // UMMA M of P2 (64 or 128)
constexpr int kTileCols = kP2M;     // columns per P2 stage
constexpr int kStageBBoxes = kTileCols / kBoxCols;
constexpr int kBoxBytes = kBlockRows * kBoxCols * 4;        // 8 KB
constexpr int kStageABoxes = 4;                             // P1 stage: 128 columns
constexpr int kStageABytes = kStageABoxes * kBoxBytes;      // 32 KB
constexpr int kStageBytes = kStageBBoxes * kBoxBytes;       // 16 / 32 KB (P2 stage)
constexpr int kHBoxBytes = kRP * kBoxCols * 4;              // 4 KB
constexpr int kWnewBytes = kRP * kBlockRows * 4;            // 8 KB  (W_b^T, K-major: [role][row])
constexpr int kWSubBytes = 16 * kRP * 4;                    // 2 KB: one warp's 16 rows of a W tile
constexpr int kMaxStagesA = 5, kMaxStagesB = 4;
constexpr float kEps = 1.1920928955078125e-07f;

// TMEM column map (512 columns allocated)
// NBUF = 2 (one CTA per block) or 3 (CTA pairs: the epilogue runs one block behind, see below)
constexpr int kColD1 = 0;                                      // NBUF x 32: XHt (M=64 layout)
__host__ __device__ constexpr int col_den(int nbuf) { return nbuf * kRP; }         // NBUF x 32: W (H H^T)
__host__ __device__ constexpr int col_wtw(int nbuf) { return 2 * nbuf * kRP; }     // 32: W^T W (rows 0..31 valid)
__host__ __device__ constexpr int col_d2(int nbuf) { return 2 * nbuf * kRP + kRP; }  // tiles x 32: (W^T X)^T
constexpr int kTmemCols = 512;
__host__ __device__ constexpr int max_tiles(int nbuf) { return (kTmemCols - col_d2(nbuf)) / kRP; }

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
    int cluster;         // CTAs per row block (1 or 2); must match the launch's cluster size
    int wide_p2;         // P2 as 8 MMAs of M = 64 (roles) x N = 256 (all columns of the CTA) per block
    int debug;           // development switches (GR_NMF_TC_DEBUG): 1 skip P1 MMAs, 2 skip P2 MMAs,
                         // 8 skip P2 TMA loads + MMAs (results are then meaningless: timing only)
    int full_n;          // development (GR_NMF_FULL_N): UMMA N = 32 whatever r is
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
// One lane of the (converged) warp.  ptxas knows a region guarded by elect.sync runs in exactly one
// thread and emits tcgen05.mma there as a plain uniform instruction; under `lane == 0` it wraps
// every MMA in an ELECT / BRA.U.ANY loop over the possibly-active lanes.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// ---- cluster helpers (CL = 2) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// 8-byte store into another CTA's shared memory whose bytes are counted on that CTA's mbarrier
__device__ __forceinline__ void st_async_v2(uint32_t addr, float a, float b, uint32_t remote_bar) {
    asm volatile(
        "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
        ::"r"(addr), "f"(a), "f"(b), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
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
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int x, int y, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(map), "r"(x), "r"(y), "r"(src) : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// wait until at most `Pending` of this thread's bulk stores still have to read shared memory
template <int Pending>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(Pending) : "memory");
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
// Same instruction with the two 64-bit shared-memory descriptors passed as 32-bit halves: the
// high words are compile-time constants and the low words differ from a per-stage base by a
// constant, so the issuing thread spends one add per operand per MMA.  (Building each descriptor
// from scratch cost 17 SASS instructions and ~105 cycles per MMA -- 4x the tensor pipe's own
// time for an M=64 N=32 K=8 tile -- and made the kernel MMA-issue-bound.)
__device__ __forceinline__ void tc_mma_tf32_split(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi,
                                                  uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
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
// 16 lanes x 32 columns of fp32 spread over all 32 threads (shape .16x256b, 4 repeats): thread t
// holds, for each 8-column group j = 0..3, v[4j + 0..1] = (lane t/4,     columns 8j + 2(t%4) + 0..1)
//                                          v[4j + 2..3] = (lane t/4 + 8, columns 8j + 2(t%4) + 0..1)
// -- the accumulator layout of an M=64 tile keeps its rows in lanes 0..15 of each lane quarter, so
// with the 32x32b shape half of the warp would sit idle in the epilogue.
__device__ __forceinline__ void tc_ld_16x32(uint32_t taddr, float (&v)[16]) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]),
          "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]),
          "=r"(u[14]), "=r"(u[15])
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
// low word: start address and leading-dimension byte offset; high word: stride byte offset,
// version, layout type
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr, uint32_t lbo) {
    return ((addr & 0x3ffff) >> 4) | ((lbo >> 4) << 16);
}
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo, uint64_t layout) {
    return (sbo >> 4) | (1u << 14) | ((uint32_t)layout << 29);
}
// instruction descriptor: D fp32 [4,6)=1, A/B tf32 [7,10)=[10,13)=2, a_major bit 15, b_major bit
// 16 (1 = MN-major), N >> 3 [17,23), M >> 4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// P1 / P1': make_idesc(64, N, 0, 0) (X K-major, H K-major); P2: make_idesc(128, N, 1, 1) (X^T and
// W_b MN-major); W^T W: make_idesc(64, N, 1, 1) -- N follows the number of roles, see the MMA warp
// wide P2: (W^T X)[roles, columns] = W_b^T (MN-major, M = 64: roles 32..63 alias 0..31) . X_b
// (MN-major, N = 256 columns = both P2 stages of the block, which are adjacent in the ring)
constexpr uint32_t kIdescP2Wide = make_idesc(64, 256, 1, 1);

// ---- shared memory carve-up ---------------------------------------------------------------------
struct SmemLayout {
    uint32_t h;        // 4 * groups boxes of 4 KB
    uint32_t ring_a;   // P1 ring: stages of 32 KB
    uint32_t ring_b;   // P2 ring: stages of 16 KB
    uint32_t wio;      // 2 (CL = 2: 3) x 8 KB: W_b tile in (TMA load, fp32, K-major SW128) / out (TMA store)
    uint32_t wnew;     // 8 KB: tf32(W_b new), [row][role] in the 32-byte-atom 128B swizzle
    uint32_t hht;      // 4 KB: H H^T as a K-major SW128 tile (TMA)
    uint32_t xch;      // CL = 2: 3 x 8 KB, the peer's partial X H^T tiles land here
    uint32_t bars;     // mbarriers
    uint32_t tmem_ptr;
    uint32_t total;
};
// groups = P1 stages per block handled by ONE CTA (ceil(all groups / cluster size))
__host__ __device__ inline SmemLayout smem_layout(int groups, int ring_a, int ring_b, int cluster) {
    SmemLayout L;
    uint32_t off = 0;
    L.h = off;      off += (uint32_t)groups * kStageABoxes * kHBoxBytes;
    off = (off + 1023u) & ~1023u;
    L.ring_a = off; off += (uint32_t)ring_a * kStageABytes;
    L.ring_b = off; off += (uint32_t)ring_b * kStageBytes;
    L.wio = off;    off += (cluster > 1 ? 3 : 2) * kWnewBytes;
    L.wnew = off;   off += kWnewBytes;
    L.hht = off;    off += kHBoxBytes;
    L.xch = off;    off += cluster > 1 ? 3 * kWnewBytes : 0;
    L.bars = off;   off += 64 * 8;
    L.tmem_ptr = off; off += 16;
    L.total = off;
    return L;
}
// barrier slots
enum { B_FULL_A = 0, B_EMPTY_A = 8, B_FULL_B = 16, B_EMPTY_B = 20, B_HFULL = 24, B_D1FULL = 25,
       B_D1EMPTY = 28, B_WFULL = 31, B_WEMPTY = 32, B_D2FULL = 33, B_WINFULL = 34, B_XCHFULL = 37,
       B_COUNT = 40 };

template <int CL>
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
    // column split: this CTA handles P1 groups [g_lo, g_lo + G) = columns [col_lo, col_hi)
    const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
    const int G0 = (p.groups + CL - 1) / CL;
    const int g_lo = rank * G0;
    const int G = min(G0, p.groups - g_lo);
    const int col_lo = g_lo * kStageABoxes * kBoxCols;
    const int col_hi = min(p.f, col_lo + G * kStageABoxes * kBoxCols);
    const int T = (col_hi - col_lo + kTileCols - 1) / kTileCols;
    const SmemLayout L = smem_layout(G0, p.ring_a, p.ring_b, CL);
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_h = s_base + L.h, s_ra = s_base + L.ring_a, s_rb = s_base + L.ring_b;
    const uint32_t s_wnew = s_base + L.wnew, s_wio = s_base + L.wio, s_hht = s_base + L.hht;
    const uint32_t s_bars = s_base + L.bars;
    volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + L.tmem_ptr);
    auto bar = [&](int slot) { return s_bars + 8u * (uint32_t)slot; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NA = p.ring_a, NB = p.ring_b;
    // D1 / Den buffers and how many blocks P2 trails P1 in the MMA issue order.  With CTA pairs
    // the epilogue is software-pipelined (the peer's partial needs ~1.5 us to cross the cluster),
    // so block i's updated W exists one block later than with a single CTA.
    constexpr int NBUF = CL > 1 ? 3 : 2, LAG = CL > 1 ? 2 : 1;   // W tile buffers: NBUF as well
    constexpr int kColDen = col_den(NBUF), kColWtW = col_wtw(NBUF), kColD2 = col_d2(NBUF);
    // row blocks of this CTA (pair): first, first + stride, ...
    const int64_t first = blockIdx.x / CL, stride = gridDim.x / CL;
    const int64_t nb = (p.n_blocks - first + stride - 1) / stride;
    const uint32_t s_xch = s_base + L.xch;

    // ---- setup ----
    if (threadIdx.x == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(bar(B_FULL_A + s), 1); mbar_init(bar(B_EMPTY_A + s), 1); }
        for (int s = 0; s < NB; ++s) { mbar_init(bar(B_FULL_B + s), 1); mbar_init(bar(B_EMPTY_B + s), 1); }
        mbar_init(bar(B_HFULL), 1);
        for (int i = 0; i < 3; ++i) {
            mbar_init(bar(B_D1FULL + i), 1);
            mbar_init(bar(B_D1EMPTY + i), 4);
            // CL = 2: one arrive.expect_tx per local epilogue warp + the peer's st.async bytes
            mbar_init(bar(B_XCHFULL + i), 4);
        }
        mbar_init(bar(B_WFULL), 4);
        mbar_init(bar(B_WEMPTY), 1);
        mbar_init(bar(B_D2FULL), 1);
        for (int i = 0; i < 3; ++i) mbar_init(bar(B_WINFULL + i), 4);
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
    // All 512 columns are allocated, so the allocation starts at lane 0 / column 0.  Using the
    // constant lets the MMA thread form TMEM addresses in uniform registers; with the address
    // loaded from shared memory every tcgen05.mma was wrapped in an ELECT / R2UR.BROADCAST loop.
    if (*tmem_ptr_s != 0) __trap();
    constexpr uint32_t tmem = 0;
    if (CL > 1) cluster_sync_all();      // the peer's barriers exist before anything arrives on them

    if (warp == 0) {
        // ================= TMA producer, P1 stream (HBM): H once, then 128-column stages =========
        if (elect_one()) {
            mbar_expect_tx(bar(B_HFULL), (uint32_t)(G * kStageABoxes + 1) * kHBoxBytes);
            for (int c = 0; c < G * kStageABoxes; ++c)
                tma_load_2d(s_h + c * kHBoxBytes, &map_h, col_lo + c * kBoxCols, 0, bar(B_HFULL));
            tma_load_2d(s_hht, &map_hht, 0, 0, bar(B_HFULL));
            uint32_t it = 0;
            for (int64_t i = 0; i < nb; ++i) {
                const int row = (int)((first + i * stride) * kBlockRows);
                GR_TRACE(0, i, 0);
                for (int g = 0; g < G; ++g, ++it) {
                    const int st = it % NA;
                    mbar_wait(bar(B_EMPTY_A + st), ((it / NA) & 1) ^ 1);
                    if (g < 7) GR_TRACE(0, i, 1 + g);
                    mbar_expect_tx(bar(B_FULL_A + st), kStageABytes);
                    for (int c = 0; c < kStageABoxes; ++c)
                        tma_load_2d(s_ra + st * kStageABytes + c * kBoxBytes,
                                    &map_x_k,
                                    col_lo + (g * kStageABoxes + c) * kBoxCols, row,
                                    bar(B_FULL_A + st));
                }
            }
        }
    } else if (warp == 6) {
        // ================= TMA producer, P2 stream (L2 re-reads, 32-byte-atom swizzle) ==========
        if (!(p.debug & 8) && elect_one()) {
            uint32_t it = 0;
            for (int64_t i = 0; i < nb; ++i) {
                const int row = (int)((first + i * stride) * kBlockRows);
                for (int t = 0; t < T; ++t, ++it) {
                    const int st = it % NB;
                    mbar_wait(bar(B_EMPTY_B + st), ((it / NB) & 1) ^ 1);
                    if (t < 8) GR_TRACE(1, i, t);
                    mbar_expect_tx(bar(B_FULL_B + st), kStageBytes);
                    for (int c = 0; c < kStageBBoxes; ++c)
                        tma_load_2d(s_rb + st * kStageBytes + c * kBoxBytes, &map_x_mn,
                                    col_lo + t * kTileCols + c * kBoxCols, row,
                                    bar(B_FULL_B + st));
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            mbar_wait(bar(B_HFULL), 0);
            // Issue order: P1(0); then P1(i), P2(i-1) for i >= 1; finally P2(last).  Measured
            // alternatives (C5, r = 32): interleaving P2(i-1) stages into P1(i) 9.6 ms, polling
            // both rings and consuming whichever stage landed 9.6 - 9.9 ms, this order 8.4 ms --
            // with 128 KB of ring space the slot turn-around bounds both streams, and blocking
            // mbarrier waits have less latency than a polling loop.
            uint32_t ita = 0, itb = 0;
            // UMMA N = the roles actually present, rounded up to the instruction's granularity
            // (8 at M = 64, 16 at M = 128): the B operands (H, H H^T, W_b) are read from shared
            // memory for every MMA, and columns r.. of every accumulator are never used.
            const int n1 = p.full_n ? kRP : (p.r + 7) & ~7, n2 = p.full_n ? kRP : (p.r + 15) & ~15;
            const uint32_t idesc_p1 = make_idesc(64, n1, 0, 0);
            const uint32_t idesc_p2 = make_idesc(kP2M, kP2M == 128 ? n2 : n1, 1, 1);
            const uint32_t idesc_wtw = make_idesc(64, n1, 1, 1);
            const int k_den = p.full_n ? 4 : (p.r + 7) >> 3;   // W_b has zero columns from r on
            for (int64_t i = 0; i < nb + LAG; ++i) {
                const bool do_p1 = i < nb, do_p2 = i >= LAG;
                const int buf1 = (int)(i % NBUF);        // D1 / Den buffer of block i
                const int wbuf = buf1;                   // W tile buffer of block i
                const int64_t j = i - LAG;
                const uint32_t d1 = tmem + kColD1 + buf1 * kRP;
                if (do_p1) {
                    // ---- P1(i): X_b . H^T into D1, 128 columns (16 k-steps) per stage
                    GR_TRACE(2, i, 0);
                    mbar_wait(bar(B_D1EMPTY + buf1), (uint32_t)(((i / NBUF) & 1) ^ 1));
                    tc_fence_after();
                    for (int g = 0; g < G; ++g, ++ita) {
                        const int st = ita % NA;
                        mbar_wait(bar(B_FULL_A + st), (ita / NA) & 1);
                        if (g == 0) GR_TRACE(2, i, 1);
                        tc_fence_after();
                        if (!(p.debug & 1)) {
                            const uint32_t a_lo = desc_lo(s_ra + st * kStageABytes, 16);
                            const uint32_t b_lo = desc_lo(s_h + g * kStageABoxes * kHBoxBytes, 16);
                            constexpr uint32_t hi = desc_hi(1024, kLayoutSw128);
#pragma unroll
                            for (int c = 0; c < kStageABoxes; ++c)
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    tc_mma_tf32_split(d1, a_lo + (c * kBoxBytes + k * 32) / 16, hi,
                                                      b_lo + (c * kHBoxBytes + k * 32) / 16, hi,
                                                      idesc_p1, (uint32_t)(g | c | k));
                        }
                        tc_commit(bar(B_EMPTY_A + st));
                    }
                    // ---- P1'(i): Den = W_b . (H H^T) from the TMA-loaded W tile.  (Its load was
                    // issued after the epilogue of block i-2, which needed P2(i-3): already done.)
                    mbar_wait(bar(B_WINFULL + wbuf), (uint32_t)((i / NBUF) & 1));
                    tc_fence_after();
                    {
                        const uint32_t a_lo = desc_lo(s_wio + wbuf * kWnewBytes, 16);
                        const uint32_t b_lo = desc_lo(s_hht, 16);
                        constexpr uint32_t hi = desc_hi(1024, kLayoutSw128);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (k < k_den)
                                tc_mma_tf32_split(tmem + kColDen + buf1 * kRP, a_lo + k * 2, hi,
                                                  b_lo + k * 2, hi, idesc_p1, (uint32_t)k);
                    }
                    tc_commit(bar(B_D1FULL + buf1));
                    GR_TRACE(2, i, 2);
                }
                if (do_p2) {
                    // ---- P2(j), j = i - LAG: (W^T X)^T and W^T W with the updated W of block j
                    mbar_wait(bar(B_WFULL), (uint32_t)(j & 1));
                    GR_TRACE(2, j, 3);
                    tc_fence_after();
                    if (p.wide_p2) {
                        // Both 128-column tiles of the block (ring stages 0 and 1: T == NB == 2, so
                        // tile t always lands in stage t) as ONE B operand of 256 columns, the tf32
                        // copy of W_b as the A operand: 8 MMAs per block instead of 16.  The tensor
                        // pipe retires these small MMAs at a fixed ~50 - 75 cycles each, so fewer
                        // and larger is what counts (profiles/README.md).
                        if (!(p.debug & 8)) {
                            mbar_wait(bar(B_FULL_B + 0), (itb / NB) & 1);
                            mbar_wait(bar(B_FULL_B + 1), ((itb + 1) / NB) & 1);
                        }
                        tc_fence_after();
                        if (!(p.debug & 10)) {
                            const uint32_t a_lo = desc_lo(s_wnew, 0);
                            const uint32_t b_lo = desc_lo(s_rb, kBoxBytes);
                            constexpr uint32_t hi = desc_hi(512, kLayoutSw128Base32);
                            const uint32_t acc0 = (uint32_t)(j != 0);
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                tc_mma_tf32_split(tmem + kColD2, a_lo + k * 64, hi, b_lo + k * 64,
                                                  hi, kIdescP2Wide, acc0 | (uint32_t)k);
                        }
                        tc_commit(bar(B_EMPTY_B + 0));
                        tc_commit(bar(B_EMPTY_B + 1));
                        itb += 2;
                    } else
                    for (int t = 0; t < T; ++t, ++itb) {
                        const int st = itb % NB;
                        if (!(p.debug & 8)) mbar_wait(bar(B_FULL_B + st), (itb / NB) & 1);
                        tc_fence_after();
                        if (!(p.debug & 10)) {
                            // k-step = 8 rows of the block = one 1 KB swizzle atom
                            const uint32_t a_lo = desc_lo(s_rb + st * kStageBytes, kBoxBytes);
                            const uint32_t b_lo = desc_lo(s_wnew, 0);
                            constexpr uint32_t hi = desc_hi(512, kLayoutSw128Base32);
                            const uint32_t acc0 = (uint32_t)(j != 0);
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                tc_mma_tf32_split(tmem + kColD2 + t * kRP, a_lo + k * 64, hi,
                                                  b_lo + k * 64, hi, idesc_p2, acc0 | (uint32_t)k);
                        }
                        tc_commit(bar(B_EMPTY_B + st));
                    }
                    // W^T W: both CTAs of a pair hold the same W_b, so they take turns (block j goes
                    // to rank j % 2) and the partials add up on the host side of the iteration
                    if (CL == 1 || (int)(j & 1) == rank) {
                        // LBO = 0: the second 32-role group of "A" aliases the first (rows 32..63 of
                        // the accumulator duplicate rows 0..31 and are never read)
                        const uint32_t d_lo = desc_lo(s_wnew, 0);
                        constexpr uint32_t hi = desc_hi(512, kLayoutSw128Base32);
                        const uint32_t acc0 = (uint32_t)(j >= CL);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            tc_mma_tf32_split(tmem + kColWtW, d_lo + k * 64, hi, d_lo + k * 64, hi,
                                              idesc_wtw, acc0 | (uint32_t)k);
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
            const int buf = (int)(i % NBUF);
            const int row = (int)((first + i * stride) * kBlockRows) + q * 16;
            mbar_expect_tx(bar(B_WINFULL + buf), kWSubBytes);
            tma_load_2d(w_sub(buf), &map_w, 0, row, bar(B_WINFULL + buf));
        };
        if (lane == 0)
            for (int b = 0; b < NBUF; ++b) load_w_sub(b);
        // Fragment coordinates (tc_ld_16x32): this thread owns rows ra and ra + 8 of the block and,
        // in every 8-column group j, the column pair 8j + cp, 8j + cp + 1.
        const int ra = q * 16 + (lane >> 2), cp = 2 * (lane & 3);
        // byte offset of (row k, even column c) in an fp32 [64][32] tile with the 128B swizzle
        // (16-byte chunk c/4 at position (c/4) ^ (k % 8)) -- the W tiles and the exchange tiles
        auto off_f32 = [](int k, int c) {
            return (uint32_t)k * 128u + (uint32_t)((((c >> 2) ^ (k & 7)) << 4) | ((c & 3) << 2));
        };
        // same for the tf32 copy read by P2 (32-byte chunk c/8 at position (c/8) ^ (k % 4))
        auto off_tf32 = [](int k, int c) {
            return (uint32_t)k * 128u + (uint32_t)((((c >> 3) ^ (k & 3)) << 5) | ((c & 7) << 2));
        };
        // W *= XHt / Den on this thread's 16 elements; the fp32 result goes back into the W tile
        // (TMA store source) and stays in wn for the tf32 copy
        auto update_w = [&](const float (&xht)[16], const float (&den)[16], float (&wn)[16],
                            unsigned char* wtile) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k = ra + 8 * h, c = 8 * j + cp, e = 4 * j + 2 * h;
                    float2* wp = reinterpret_cast<float2*>(wtile + off_f32(k, c));
                    const float2 w = *wp;
                    const float d0 = den[e] == 0.f ? kEps : den[e];
                    const float d1 = den[e + 1] == 0.f ? kEps : den[e + 1];
                    // The updated W is rounded to tf32 (nearest) BEFORE it is stored: next
                    // iteration's W (H H^T) MMA reads this tile as raw fp32 bits, i.e. truncated to
                    // tf32, and a truncated operand is low by ~3.4e-4 on average -- the denominator
                    // was biased, W grew and H shrank by that factor every iteration (measured
                    // against sklearn from a shared start: both factors off by 1.8e-2 after 50
                    // iterations with the reconstruction equal to 1e-6).  With W held at tf32
                    // precision the truncation is exact and every operand rounding is unbiased.
                    wn[e] = c < r ? to_tf32(w.x * __fdividef(xht[e], d0)) : 0.f;
                    wn[e + 1] = c + 1 < r ? to_tf32(w.y * __fdividef(xht[e + 1], d1)) : 0.f;
                    *wp = make_float2(wn[e], wn[e + 1]);
                }
        };
        auto store_tf32 = [&](const float (&wn)[16]) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k = ra + 8 * h, c = 8 * j + cp, e = 4 * j + 2 * h;
                    // wn is already rounded to tf32 by update_w
                    *reinterpret_cast<float2*>(smem + L.wnew + off_tf32(k, c)) =
                        make_float2(wn[e], wn[e + 1]);
                }
        };
        // W tile out (rows beyond n / roles beyond r are clipped by the tensor map), then refill a
        // free tile buffer.  One CTA per block: two buffers, wait until this store has read its
        // tile and load block i + 2 into it.  CTA pair: three buffers, so only the PREVIOUS
        // block's store has to be done -- its buffer takes block i + 2 and the wait is off the
        // epilogue's critical path.
        auto store_w_and_refill = [&](int64_t i, int wb, bool do_store) {
            const int row = (int)((first + i * stride) * kBlockRows) + q * 16;
            if (do_store) {
                tma_store_2d(&map_w, 0, row, w_sub(wb));
                tma_store_commit();
                if constexpr (CL == 1) tma_store_wait_read<0>();
            }
            // CTA pair: the refill is issued at the START of the next epilogue iteration (below)
            if constexpr (CL == 1) load_w_sub(i + 2);
        };
        if constexpr (CL == 1) {
            for (int64_t i = 0; i < nb; ++i) {
                const int buf = (int)(i & 1);
                const uint32_t par = (uint32_t)((i >> 1) & 1);
                if (q == 0 && lane == 0) GR_TRACE(3, i, 0);
                // ---- XHt and Den from TMEM (M=64 layout: rows 16q .. 16q+15 in lanes 0..15 of
                // lane quarter q)
                mbar_wait(bar(B_D1FULL + buf), par);
                mbar_wait(bar(B_WINFULL + buf), par);        // the W tile the MMA already consumed
                tc_fence_after();
                if (q == 0 && lane == 0) GR_TRACE(3, i, 1);
                float xht[16], den[16], wn[16];
                tc_ld_16x32(tmem + lane_base + kColD1 + buf * kRP, xht);
                tc_ld_16x32(tmem + lane_base + kColDen + buf * kRP, den);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(B_D1EMPTY + buf));
                if (q == 0 && lane == 0) GR_TRACE(3, i, 2);
                update_w(xht, den, wn, smem + L.wio + buf * kWnewBytes);
                if (q == 0 && lane == 0) GR_TRACE(3, i, 3);
                mbar_wait(bar(B_WEMPTY), (uint32_t)((i & 1) ^ 1));   // P2(i-1) is done with wnew
                if (q == 0 && lane == 0) GR_TRACE(3, i, 4);
                store_tf32(wn);
                if (q == 0 && lane == 0) GR_TRACE(3, i, 5);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bar(B_WFULL));
                    store_w_and_refill(i, buf, true);
                }
                __syncwarp();
                if (q == 0 && lane == 0) GR_TRACE(3, i, 6);
            }
        } else {
            // ---- CTA pair: X H^T is a sum over columns, each CTA holds the partial of its half.
            // Software pipeline, iteration j:
            //   (1) the peer's partial of block j-1 (sent one iteration ago) -> registers, re-arm
            //       that exchange buffer for block j+2;
            //   (2) own partial of block j: TMEM -> the peer's exchange buffer (st.async, the
            //       bytes themselves complete the peer's mbarrier: one hop, no fence);
            //   (3) W update of block j-1 from own partial + peer's partial (a + b == b + a bit
            //       for bit, so both CTAs get the same W), tf32 copy, W tile store (rank 0).
            // Step (1) precedes step (2) in program order, which is what makes three exchange
            // buffers enough: the peer can only write block j+2 after it has used this CTA's
            // partial of block j, sent in step (2).
            const uint32_t peer = (uint32_t)(rank ^ 1);
            if (lane == 0)
                for (int b = 0; b < 3 && b < nb; ++b) mbar_expect_tx(bar(B_XCHFULL + b), kWSubBytes);
            float pr[16];
            for (int64_t jj = 0; jj <= nb; ++jj) {
                const int64_t i = jj - 1;                    // block of steps (1) and (3)
                if (q == 0 && lane == 0 && jj < nb) GR_TRACE(3, jj, 0);
                // Refill the W tile buffer of block i - 1 (its store was issued at the end of the
                // previous iteration) with block i + 2 BEFORE this iteration's waits and its W
                // update: issued after the update, the tile of block i + 2 trailed WFULL(i) by a
                // TMA round trip and the MMA thread spun on WINFULL in every block -- the W tile
                // load sat on the loop-carried chain load -> P1' -> D1FULL -> exchange -> update
                // -> load (the kernel took 4.0 ms per C5 iteration even with every MMA skipped).
                if (i >= 1 && lane == 0) {
                    tma_store_wait_read<0>();
                    load_w_sub(i + 2);
                }
                if (i >= 0) {
                    const int b = (int)(i % 3);
                    mbar_wait_acquire_cluster(bar(B_XCHFULL + b), (uint32_t)((i / 3) & 1));
                    const unsigned char* xt = smem + L.xch + b * kWnewBytes;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const float2 t = *reinterpret_cast<const float2*>(
                                xt + off_f32(ra + 8 * h, 8 * j + cp));
                            pr[4 * j + 2 * h] = t.x;
                            pr[4 * j + 2 * h + 1] = t.y;
                        }
                    __syncwarp();
                    if (lane == 0 && i + 3 < nb) mbar_expect_tx(bar(B_XCHFULL + b), kWSubBytes);
                }
                if (jj < nb) {
                    const int b = (int)(jj % 3);
                    mbar_wait(bar(B_D1FULL + b), (uint32_t)((jj / 3) & 1));
                    tc_fence_after();
                    if (q == 0 && lane == 0) GR_TRACE(3, jj, 1);
                    float part[16];
                    tc_ld_16x32(tmem + lane_base + kColD1 + b * kRP, part);
                    const uint32_t dst = mapa(s_xch + b * kWnewBytes, peer);
                    const uint32_t rbar = mapa(bar(B_XCHFULL + b), peer);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            st_async_v2(dst + off_f32(ra + 8 * h, 8 * j + cp), part[4 * j + 2 * h],
                                        part[4 * j + 2 * h + 1], rbar);
                    if (q == 0 && lane == 0) GR_TRACE(3, jj, 2);
                }
                if (i >= 0) {
                    const int b = (int)(i % 3), wb = b;
                    mbar_wait(bar(B_WINFULL + wb), (uint32_t)((i / 3) & 1));    // W tile of block i
                    float xht[16], den[16], wn[16];
                    tc_ld_16x32(tmem + lane_base + kColD1 + b * kRP, xht);
                    tc_ld_16x32(tmem + lane_base + kColDen + b * kRP, den);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(B_D1EMPTY + b));
#pragma unroll
                    for (int l = 0; l < 16; ++l) xht[l] += pr[l];
                    update_w(xht, den, wn, smem + L.wio + wb * kWnewBytes);
                    if (q == 0 && lane == 0) GR_TRACE(3, i, 3);
                    mbar_wait(bar(B_WEMPTY), (uint32_t)((i & 1) ^ 1));   // P2(i-1) is done with wnew
                    if (q == 0 && lane == 0) GR_TRACE(3, i, 4);
                    store_tf32(wn);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(bar(B_WFULL));
                        store_w_and_refill(i, wb, rank == 0);
                    }
                    __syncwarp();
                    if (q == 0 && lane == 0) GR_TRACE(3, i, 6);
                }
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");

        // ---- final: dump the TMEM accumulators as this CTA's partials
        mbar_wait(bar(B_D2FULL), 0);
        tc_fence_after();
        float v[32];
        if (p.wide_p2) {
            // M = 64 layout: role 16q + l in lane l < 16 of quarter q (q < 2), TMEM column = the
            // CTA's column; every thread writes 32 consecutive columns of its role
            for (int ch = 0; ch < 8; ++ch) {
                tc_ld_32x32(tmem + lane_base + kColD2 + ch * 32, v);
                const int role = q * 16 + lane;
                if (q < 2 && lane < 16 && role < r)
#pragma unroll
                    for (int l = 0; l < 32; ++l) {
                        const int col = col_lo + ch * 32 + l;
                        if (col < col_hi)
                            p.part_wtx[((int64_t)blockIdx.x * kRP + role) * p.f + col] = v[l];
                    }
            }
        } else
        for (int t = 0; t < T; ++t) {
            tc_ld_32x32(tmem + lane_base + kColD2 + t * kRP, v);
            // M=64 layout: row 16q + l in lane l < 16 of quarter q; M=128: row 32q + l in lane l
            const int col = col_lo + t * kTileCols + (kP2M == 64 ? q * 16 : q * 32) + lane;
            if ((kP2M == 128 || lane < 16) && col < col_hi)
#pragma unroll
                for (int l = 0; l < 32; ++l)
                    if (l < r)
                        p.part_wtx[((int64_t)blockIdx.x * kRP + l) * p.f + col] = v[l];
        }
        tc_ld_32x32(tmem + lane_base + kColWtW, v);
        if (nb > rank && q < 2 && lane < 16) {      // this CTA accumulated at least one block
            const int role = q * 16 + lane;   // M=64 layout
            if (role < r)
#pragma unroll
                for (int l = 0; l < 32; ++l)
                    if (l < r) p.part_wtw[((int64_t)blockIdx.x * kRP + role) * r + l] = v[l];
        }
        tc_fence_before();
    }

    __syncthreads();
    if (CL > 1) cluster_sync_all();      // no remote store / arrive may target a CTA that has left
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

// ---- convergence check on the tensor core: ||X - W H||_F^2 ------------------------------------
// sklearn evaluates the dense residual every 10 iterations (_nmf.py:867-879 -> :122).  On CUDA
// cores that pass is bound by FFMA issue (n f r FMAs: 10.8 ms on C5 at r = 32 against 3.5 ms for
// streaming X once, nmf_mu.cu).  Here W_b H runs as tcgen05.mma (M = 128 rows, N = 64 columns,
// K = 8 roles x ceil(r / 8)) into TMEM, and the epilogue warps only subtract and square:
//   warp 0     TMA producer: H^T of the CTA's 256-column span once (B operand, K-major), per
//              128-row block the W tile (A operand, K-major), per stage two [128 x 32] boxes of X
//              (plain fp32: X never passes through the tensor core)
//   warp 1     MMA issuer: one accumulator of 64 TMEM columns per stage, four in rotation
//   warps 2-9  epilogue: lane quarter warp % 4, column half (warp - 2) / 4; thread = one row x 32
//              columns: tcgen05.ld.32x32b, eight conflict-free LDS.128 of the swizzled X box,
//              fp32 partial per stage (32 terms), fp64 running total per thread
// W and H^T are loaded with the TFLOAT32 element type (round to nearest; W is already held at
// tf32 precision by the iteration kernel), so the only perturbation is the rounding of H: the
// error moves by ~1e-7 relative on the C5 data (tests/test_nmf_gpu.py states 1e-5).
// CTA c works on column span c % spans of the row blocks c / spans, c / spans + grid / spans, ...
// Partials: out[8 * CTA + epilogue warp], summed by the host in that order.
constexpr int kErrTcThreads = 320;
constexpr int kErrTcRows = 128;                              // rows per block = UMMA M
constexpr int kErrTcChunk = 64;                              // columns per stage = UMMA N
constexpr int kErrTcSpan = 256;                              // columns per CTA
constexpr int kErrTcBoxBytes = kErrTcRows * kBoxCols * 4;    // 16 KB: [128 rows x 32 columns] of X
constexpr int kErrTcStageBytes = 2 * kErrTcBoxBytes;         // 32 KB
constexpr int kErrTcWBytes = kErrTcRows * kRP * 4;           // 16 KB
constexpr int kErrTcHtBytes = kErrTcSpan * kRP * 4;          // 32 KB
constexpr int kErrTcStages = 5, kErrTcDBufs = 4;
constexpr uint32_t kErrTcSmem = kErrTcHtBytes + 2 * kErrTcWBytes + kErrTcStages * kErrTcStageBytes +
                                64 * 8 + 1024;
enum { E_FULL = 0, E_EMPTY = 5, E_HFULL = 10, E_WFULL = 11, E_WEMPTY = 13, E_DFULL = 15,
       E_DEMPTY = 19, E_COUNT = 23 };

struct ErrTcParams {
    int64_t n_blocks;    // ceil(n / 128)
    int f, r;
    int spans;           // ceil(f / 256)
    double* out;         // [grid, 8]
};

__global__ void nmf_transpose_h_kernel(const float* __restrict__ H, int r, int f,
                                       float* __restrict__ Ht) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;     // Ht[col][role], 32 roles per column
    if (i >= f * kRP) return;
    const int col = i / kRP, role = i % kRP;
    Ht[i] = role < r ? H[(int64_t)role * f + col] : 0.f;
}

__global__ void __launch_bounds__(kErrTcThreads, 1)
nmf_error_tc_kernel(const __grid_constant__ CUtensorMap map_x,
                    const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_ht, const ErrTcParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t s_ht = smem_u32(smem), s_w = s_ht + kErrTcHtBytes;
    const uint32_t s_ring = s_w + 2 * kErrTcWBytes;
    const uint32_t s_bars = s_ring + kErrTcStages * kErrTcStageBytes;
    auto bar = [&](int slot) { return s_bars + 8u * (uint32_t)slot; };
    __shared__ uint32_t tmem_ptr_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int span = (int)blockIdx.x % p.spans;
    const int64_t first = blockIdx.x / p.spans, stride = gridDim.x / p.spans;
    const int64_t nb = (p.n_blocks - first + stride - 1) / stride;
    const int col_lo = span * kErrTcSpan;
    const int cols = min(p.f - col_lo, kErrTcSpan);
    const int nch = (cols + kErrTcChunk - 1) / kErrTcChunk;      // stages per row block
    const int nboxes = (cols + kBoxCols - 1) / kBoxCols;         // 32-column boxes that hold data

    if (threadIdx.x == 0) {
        for (int s = 0; s < kErrTcStages; ++s) {
            mbar_init(bar(E_FULL + s), 1);
            mbar_init(bar(E_EMPTY + s), 8);
        }
        mbar_init(bar(E_HFULL), 1);
        for (int b = 0; b < 2; ++b) { mbar_init(bar(E_WFULL + b), 1); mbar_init(bar(E_WEMPTY + b), 1); }
        for (int b = 0; b < kErrTcDBufs; ++b) {
            mbar_init(bar(E_DFULL + b), 1);
            mbar_init(bar(E_DEMPTY + b), 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_ptr_s)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tmem_ptr_s != 0) __trap();      // all 512 columns: the allocation starts at 0
    constexpr uint32_t tmem = 0;

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(bar(E_HFULL), (uint32_t)nboxes * kHBoxBytes);
            for (int b = 0; b < nboxes; ++b)
                tma_load_2d(s_ht + b * kHBoxBytes, &map_ht, 0, col_lo + b * kBoxCols, bar(E_HFULL));
            uint32_t it = 0;
            for (int64_t i = 0; i < nb; ++i) {
                const int row = (int)((first + i * stride) * kErrTcRows);
                const int wb = (int)(i & 1);
                mbar_wait(bar(E_WEMPTY + wb), (uint32_t)(((i >> 1) & 1) ^ 1));
                mbar_expect_tx(bar(E_WFULL + wb), kErrTcWBytes);
                tma_load_2d(s_w + wb * kErrTcWBytes, &map_w, 0, row, bar(E_WFULL + wb));
                for (int c = 0; c < nch; ++c, ++it) {
                    const int st = it % kErrTcStages;
                    mbar_wait(bar(E_EMPTY + st), ((it / kErrTcStages) & 1) ^ 1);
                    const int nh = min(2, nboxes - 2 * c);
                    mbar_expect_tx(bar(E_FULL + st), (uint32_t)nh * kErrTcBoxBytes);
                    for (int hf = 0; hf < nh; ++hf)
                        tma_load_2d(s_ring + st * kErrTcStageBytes + hf * kErrTcBoxBytes, &map_x,
                                    col_lo + c * kErrTcChunk + hf * kBoxCols, row,
                                    bar(E_FULL + st));
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            mbar_wait(bar(E_HFULL), 0);
            tc_fence_after();
            const int ks = (p.r + 7) >> 3;          // W and H^T hold zeros from role r on
            constexpr uint32_t idesc = make_idesc(kErrTcRows, kErrTcChunk, 0, 0);
            constexpr uint32_t hi = desc_hi(1024, kLayoutSw128);
            uint32_t it = 0;
            for (int64_t i = 0; i < nb; ++i) {
                const int wb = (int)(i & 1);
                mbar_wait(bar(E_WFULL + wb), (uint32_t)((i >> 1) & 1));
                tc_fence_after();
                const uint32_t a_lo = desc_lo(s_w + wb * kErrTcWBytes, 16);
                for (int c = 0; c < nch; ++c, ++it) {
                    const int db = it % kErrTcDBufs;
                    mbar_wait(bar(E_DEMPTY + db), ((it / kErrTcDBufs) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t b_lo = desc_lo(s_ht + c * 2 * kHBoxBytes, 16);
                    for (int k = 0; k < ks; ++k)
                        tc_mma_tf32_split(tmem + db * kErrTcChunk, a_lo + k * 2, hi, b_lo + k * 2,
                                          hi, idesc, (uint32_t)k);
                    tc_commit(bar(E_DFULL + db));
                }
                tc_commit(bar(E_WEMPTY + wb));
            }
        }
    } else {
        const int q = warp & 3, hf = (warp - 2) >> 2;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int k = q * 32 + lane;                 // this thread's row of the block
        double total = 0.0;
        uint32_t it = 0;
        for (int64_t i = 0; i < nb; ++i)
            for (int c = 0; c < nch; ++c, ++it) {
                const int st = it % kErrTcStages, db = it % kErrTcDBufs;
                // Both waits also in a warp whose column half holds no data (ragged span): a warp
                // that ran ahead would arrive twice within one phase of the EMPTY barriers.
                mbar_wait(bar(E_DFULL + db), (it / kErrTcDBufs) & 1);
                mbar_wait(bar(E_FULL + st), (it / kErrTcStages) & 1);
                if (hf < min(2, nboxes - 2 * c)) {
                    tc_fence_after();
                    float wh[32];
                    tc_ld_32x32(tmem + lane_base + db * kErrTcChunk + hf * kBoxCols, wh);
                    tc_fence_before();
                    // row k of the box: 128 bytes, 16-byte chunk j at position j ^ (k % 8)
                    const uint32_t xr = s_ring + st * kErrTcStageBytes + hf * kErrTcBoxBytes + k * 128;
                    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 x;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                                     : "r"(xr + ((j ^ (k & 7)) << 4)));
                        const float d0 = x.x - wh[4 * j], d1 = x.y - wh[4 * j + 1];
                        const float d2 = x.z - wh[4 * j + 2], d3 = x.w - wh[4 * j + 3];
                        p0 = fmaf(d0, d0, p0);
                        p1 = fmaf(d1, d1, p1);
                        p2 = fmaf(d2, d2, p2);
                        p3 = fmaf(d3, d3, p3);
                    }
                    total += (double)((p0 + p1) + (p2 + p3));
                }
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bar(E_DEMPTY + db));
                    mbar_arrive(bar(E_EMPTY + st));
                }
            }
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        if (lane == 0) p.out[(int64_t)blockIdx.x * 8 + (warp - 2)] = total;
    }

    tc_fence_before();
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
    int cluster = 1;
    int ring_a = 0, ring_b = 0;
    size_t smem_bytes = 0;
    float* d_part_wtx = nullptr;
    float* d_part_wtw = nullptr;
    CUtensorMap map_x_k, map_x_mn, map_h, map_hht, map_w;
    const float* W = nullptr;
    const float* X = nullptr;   // what the X maps were encoded for
    int64_t ldx = 0;
    const float* H = nullptr;
    // convergence check (nmf_error_tc)
    CUtensorMap map_ex, map_ew, map_eht;
    float* d_ht = nullptr;      // [f, 32]: H transposed, zero from role r on
    const float* eX = nullptr;
    int64_t eldx = 0;
    const float* eW = nullptr;
    bool err_ready = false;
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

// CTAs per row block for f columns: 2 (column-split pair) when every CTA gets at least one
// 128-column group and its tiles fit TMEM beside three D1 / Den buffers, else 1; 0 = the
// accumulators do not fit TMEM at all.
constexpr size_t kSmemOptin = 232448;   // sm_100: 227 KB of dynamic shared memory per CTA
static bool smem_fits(int groups_cta, int ring_a, int ring_b, int cluster) {
    return (size_t)smem_layout(groups_cta, ring_a, ring_b, cluster).total + 1024 <= kSmemOptin;
}
static int pick_cluster(int f) {
    const int groups = ceil_div(f, kStageABoxes * kBoxCols);
    const int cols_cta = ceil_div(groups, 2) * kStageABoxes * kBoxCols;
    if (groups >= 2 && !getenv("GR_NMF_NO_CLUSTER") &&
        ceil_div(std::min(f, cols_cta), kTileCols) <= max_tiles(3) &&
        smem_fits(ceil_div(groups, 2), 2, 1, 2))
        return 2;
    return ceil_div(f, kTileCols) <= max_tiles(2) && smem_fits(groups, 2, 1, 1) ? 1 : 0;
}

bool gr::nmf_tc_supported(const gr_nmf* h, const float* X, int64_t ldx) {
    if (getenv("GR_NMF_DISABLE_TC")) return false;
    return h->r <= kRP && h->r % 4 == 0 && h->f % 4 == 0 && ldx % 4 == 0 && aligned16(X) &&
           pick_cluster(h->f) != 0 && h->n < ((int64_t)1 << 31) &&
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
    }
    if (!s->grid) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
        const int64_t n_blocks = ceil_div<int64_t>(h->n, kBlockRows);
        // column-split CTA pairs need at least one 128-column group per CTA
        s->cluster = pick_cluster(h->f);
        s->grid = s->cluster * (int)std::min<int64_t>(sms / s->cluster, n_blocks);
        const int groups_cta = ceil_div(groups, s->cluster);
        int max_smem = 0;
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
        auto fits = [&](int a, int b) {
            return (size_t)smem_layout(groups_cta, a, b, s->cluster).total + 1024 <=
                   (size_t)max_smem;
        };
        // One CTA per block: P1 (HBM) ring of three 32 KB stages (they stream at > 6.5 TB/s,
        // tools/exp_tma.cu), P2 (L2) ring of two 16 KB stages, what is left goes to the P2 ring
        // first (its slots turn around faster), then to the P1 ring.
        // CTA pairs: a P1 ring of two stages already holds a whole half-block at f = 512, and
        // a P2 ring that holds all four tiles of a half-block is worth more than a third P1
        // stage (measured on C5: rings 2/4 5.2 ms, 3/2 6.4 ms per iteration).
        s->ring_a = s->cluster == 2 ? 2 : 3;
        s->ring_b = 2;
        if (!fits(s->ring_a, s->ring_b)) s->ring_a = 2;
        if (!fits(s->ring_a, s->ring_b)) s->ring_b = 1;
        if (!fits(s->ring_a, s->ring_b)) return fail(GR_ERR_CUDA, "nmf tc: shared memory budget");
        while (s->ring_b < kMaxStagesB && fits(s->ring_a, s->ring_b + 1)) ++s->ring_b;
        while (s->ring_a < kMaxStagesA && fits(s->ring_a + 1, s->ring_b)) ++s->ring_a;
        if (getenv("GR_NMF_RING_A") && getenv("GR_NMF_RING_B") &&
            fits(atoi(getenv("GR_NMF_RING_A")), atoi(getenv("GR_NMF_RING_B")))) {
            s->ring_a = std::max(1, std::min(kMaxStagesA, atoi(getenv("GR_NMF_RING_A"))));
            s->ring_b = std::max(1, std::min(kMaxStagesB, atoi(getenv("GR_NMF_RING_B"))));
        }
        s->smem_bytes = (size_t)smem_layout(groups_cta, s->ring_a, s->ring_b, s->cluster).total + 1024;
        if (s->cluster == 2)
            GR_CUDA_TRY(cudaFuncSetAttribute(nmf_fused_tc_kernel<2>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)s->smem_bytes));
        else
            GR_CUDA_TRY(cudaFuncSetAttribute(nmf_fused_tc_kernel<1>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)s->smem_bytes));
        const size_t wtx_bytes = (size_t)s->grid * kRP * h->f * sizeof(float);
        const size_t wtw_bytes = (size_t)s->grid * kRP * h->r * sizeof(float);
        GR_CUDA_TRY(cudaMalloc(&s->d_part_wtx, wtx_bytes));
        GR_CUDA_TRY(cudaMalloc(&s->d_part_wtw, wtw_bytes));
        // a CTA of a pair writes only its own columns of W^T X (and rank 0 alone W^T W): the
        // rest of its partial block stays zero
        GR_CUDA_TRY(cudaMemsetAsync(s->d_part_wtx, 0, wtx_bytes, st));
        GR_CUDA_TRY(cudaMemsetAsync(s->d_part_wtw, 0, wtw_bytes, st));
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
    p.cluster = s->cluster;
    p.full_n = getenv("GR_NMF_FULL_N") != nullptr;
    // wide P2 needs two 128-column tiles per CTA in a two-stage ring (so that a block's tiles sit
    // side by side in shared memory) and 256 TMEM columns behind the D1 / Den / W^T W buffers
    // Opt-in (GR_NMF_WIDE_P2=1): measured on C5 after the early W refill, 16 MMAs of M = 128 x
    // N = 32 (4.14 / 4.29 / 4.33 / 4.36 ms for r = 4 / 8 / 16 / 32) beat 8 of M = 64 x N = 256
    // (4.24 / 4.33 / 4.37 / 4.62 ms): an M = 64 MMA runs the tensor pipe at half rate.
    p.wide_p2 = s->cluster == 2 && groups == 4 && s->ring_b == 2 && kP2M == 128 &&
                col_d2(3) + 256 <= kTmemCols && getenv("GR_NMF_WIDE_P2") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)s->grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = s->smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)s->cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = s->cluster == 2
        ? cudaLaunchKernelEx(&cfg, nmf_fused_tc_kernel<2>, s->map_x_k, s->map_x_mn, s->map_h,
                             s->map_hht, s->map_w, p)
        : cudaLaunchKernelEx(&cfg, nmf_fused_tc_kernel<1>, s->map_x_k, s->map_x_mn, s->map_h,
                             s->map_hht, s->map_w, p);
    count_launch();
    if (e == cudaSuccess) e = cudaGetLastError();
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

// ||X - W H||_F on the tensor core (nmf_error_tc_kernel); same shapes as nmf_tc_supported.
int gr::nmf_error_tc(gr_nmf* h, const float* X, int64_t ldx, const float* W, const float* H,
                     double* err, cudaStream_t st) {
    TcState* s = static_cast<TcState*>(h->tc_state);
    if (!s) {
        s = new (std::nothrow) TcState();
        if (!s) return fail(GR_ERR_OUT_OF_MEMORY, "nmf tc state");
        h->tc_state = s;
    }
    if (!s->err_ready) {
        GR_CUDA_TRY(cudaMalloc(&s->d_ht, (size_t)h->f * kRP * sizeof(float)));
        GR_CUDA_TRY(cudaFuncSetAttribute(nmf_error_tc_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kErrTcSmem));
        if (int rc = encode_2d(&s->map_eht, s->d_ht, (uint64_t)kRP, (uint64_t)h->f,
                               (uint64_t)kRP * 4, kRP, kBoxCols, CU_TENSOR_MAP_SWIZZLE_128B))
            return rc;
        s->err_ready = true;
    }
    if (s->eX != X || s->eldx != ldx) {
        // plain fp32: X is read by the epilogue threads, not by the tensor core
        if (int rc = encode_2d(&s->map_ex, X, (uint64_t)h->f, (uint64_t)h->n, (uint64_t)ldx * 4,
                               kBoxCols, kErrTcRows, CU_TENSOR_MAP_SWIZZLE_128B, true))
            return rc;
        s->eX = X;
        s->eldx = ldx;
    }
    if (s->eW != W) {
        if (int rc = encode_2d(&s->map_ew, W, (uint64_t)h->r, (uint64_t)h->n, (uint64_t)h->r * 4,
                               kRP, kErrTcRows, CU_TENSOR_MAP_SWIZZLE_128B))
            return rc;
        s->eW = W;
    }
    nmf_transpose_h_kernel<<<ceil_div(h->f * kRP, 256), 256, 0, st>>>(H, h->r, h->f, s->d_ht);
    count_launch();
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    ErrTcParams p;
    p.n_blocks = ceil_div<int64_t>(h->n, kErrTcRows);
    p.f = h->f;
    p.r = h->r;
    p.spans = ceil_div(h->f, kErrTcSpan);
    p.out = h->d_err_part;
    const int streams = (int)std::min<int64_t>(std::max(1, sms / p.spans), p.n_blocks);
    const int grid = streams * p.spans;
    nmf_error_tc_kernel<<<grid, kErrTcThreads, kErrTcSmem, st>>>(s->map_ex, s->map_ew, s->map_eht,
                                                                p);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return fail(GR_ERR_CUDA, "nmf_error_tc_kernel launch failed: %s", cudaGetErrorString(e));
    const size_t cnt = (size_t)grid * 8;
    h->h_err_part.resize(cnt);
    GR_CUDA_TRY(cudaMemcpyAsync(h->h_err_part.data(), h->d_err_part, cnt * sizeof(double),
                                cudaMemcpyDeviceToHost, st));
    GR_CUDA_TRY(cudaStreamSynchronize(st));
    double total = 0.0;
    for (size_t i = 0; i < cnt; ++i) total += h->h_err_part[i];  // fixed order
    *err = std::sqrt(total);
    return GR_OK;
}

void gr::nmf_tc_release(gr_nmf* h) {
    TcState* s = static_cast<TcState*>(h->tc_state);
    if (!s) return;
    cudaFree(s->d_ht);
    cudaFree(s->d_part_wtx);
    cudaFree(s->d_part_wtw);
    delete s;
    h->tc_state = nullptr;
}
