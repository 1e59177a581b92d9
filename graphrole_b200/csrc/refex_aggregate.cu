// Path A: one ReFeX recursion level as a CSR gather-reduce (fused sum + mean) for sm_100a.
//
// Replaces the per-node pandas chain of RecursiveFeatureExtractor._get_next_features
// (graphrole/features/extract.py:105-118): reindex(neighbours) -> agg([sum, mean]) ->
// fillna(0), i.e. S = A.X and M = S / outdeg with A the binary out-adjacency.
//
// The kernel is HBM bound (0.25 flop/byte): everything below is about keeping many 16-byte
// row-gather requests in flight per SM and not moving a byte twice.
//   * a warp owns `rows_per_warp` CONSECUTIVE rows, so its colidx span is one contiguous
//     stream, read in 128-byte-aligned chunks of 32 indices with one chunk of prefetch, and
//     its rowptr entries are one coalesced load; only the X-row gathers are dependent loads;
//   * a feature row (d floats) is covered by LPR = d/4 lanes with one float4 each and the
//     32/LPR lane groups of the warp take different neighbours.  Measured on B200
//     (profiles/README.md): the kernel is bound by the number of warps with a gather in
//     flight, not by loads per warp -- 32 registers / 64 resident warps per SM with ONE
//     gather per lane beats every unrolled variant (23.8 -> 17.4 ms per level on C3);
//   * colidx and the outputs are touched once and use streaming (evict-first) accesses, the
//     gathers bypass L1 (no reuse there), and rows the handle tagged as hot (sign bit of the
//     library's colidx copy, GR_CSR_HOT_HINTS) are loaded with an L2 evict_last policy so
//     the rows power-law graphs keep coming back to stay resident in L2;
//   * rows longer than kHubThreshold arcs are cut into kHubSegment-arc segments handled by
//     the leading CTAs of the same launch; a small second kernel adds a row's partials in
//     fp64 in segment order (bitwise reproducible, no float atomics).

#include <algorithm>

#include "csr_handle.cuh"

using namespace gr;

namespace {

constexpr int kWarps = 8;      // 256 threads per CTA
constexpr int kMinBlocks = 8;  // 32 registers/thread -> 64 resident warps per SM
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxReplicas = 16;
// Gathers per lane in flight for rows of 64 / 128 bytes (LPR = 4 / 8).  Measured on C3
// (profiles/r2_narrow_u.txt): d = 32 all rows 10.39 -> 9.62 ms, a hub-heavy quarter 4.50 -> 3.06,
// a tail quarter 2.88 -> 2.75; d = 16 7.74 -> 7.55; 32-byte rows (LPR = 2) do not gain.
constexpr int kNarrowRowGathers = 2;

struct RefexArgs {
    const int64_t* __restrict__ rowptr;
    const int32_t* __restrict__ colidx;
    const float* __restrict__ X;
    int64_t ldx;
    int32_t d;
    int64_t row_lo, row_hi;
    float* __restrict__ out_sum;
    float* __restrict__ out_mean;
    int64_t ldo;
    const int64_t* __restrict__ seg_begin;
    const int64_t* __restrict__ seg_end;
    int64_t seg_lo, seg_hi;
    float* __restrict__ partial;  // [n_segments, d], indexed by handle-global segment id
    int64_t n_seg_blocks;         // leading CTAs (blockIdx.x) that reduce hub segments
    int32_t rows_per_warp;
    uint32_t hub_threshold;       // rows with more arcs are left to the segment warps
};

// Fused gather + broadcast (node-range sharded recursion): the mean rows are written into
// `n_rep` replicas of the next level's input matrix -- this GPU's own and its peers', the latter
// mapped over NVLink -- instead of one local buffer, so the exchange step of the level rides on
// the gather kernel's own stores.  Pointers are rebased to handle-local row 0.
struct Replicas {
    float* mean[kMaxReplicas];
    int32_t n_rep;
};

// ---- vector helpers -------------------------------------------------------------------
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// Gather of one lane's slice of a feature row: L1 bypassed (no reuse there); rows tagged hot
// carry an L2 evict_last policy.  Cold rows use the plain form: measured on B200, attaching an
// explicit evict_normal policy to every cold load costs 6 % (18.7 vs 17.6 ms per C3 level),
// more than the occasional hot/cold divergence between lane groups.
template <int VW>
__device__ __forceinline__ void load_row(float (&v)[VW], const float* p, bool ok, bool hot,
                                         uint64_t policy);

template <>
__device__ __forceinline__ void load_row<4>(float (&v)[4], const float* p, bool ok, bool hot,
                                            uint64_t policy) {
    v[0] = v[1] = v[2] = v[3] = 0.f;
    if (ok) {
        if (hot)
            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                         : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p), "l"(policy));
        else
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
    }
}
template <>
__device__ __forceinline__ void load_row<1>(float (&v)[1], const float* p, bool ok, bool hot,
                                            uint64_t policy) {
    v[0] = 0.f;
    if (ok) {
        if (hot)
            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;"
                         : "=f"(v[0]) : "l"(p), "l"(policy));
        else
            asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v[0]) : "l"(p));
    }
}

// The mean is the correctly rounded fp32 quotient sum / count (__fdiv_rn), never sum * (1 / count):
// the reference divides in float64 (pandas mean), which keeps exact ties -- seven neighbours that
// all carry 3.0 have mean exactly 3.0 -- and vertical_log_binning (prune.py:27-45) is rank based,
// so a tie broken by a reciprocal's rounding error moves bin boundaries.  `count` is max(deg, 1)
// (an empty row has sum 0).
template <int VW>
__device__ __forceinline__ void store_sum(float* p, const float (&v)[VW]) {
    if constexpr (VW == 4)
        __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    else
        __stcs(p, v[0]);
}

template <int VW>
__device__ __forceinline__ void store_stream(float* p, const float (&v)[VW], float count);

template <>
__device__ __forceinline__ void store_stream<4>(float* p, const float (&v)[4], float count) {
    __stcs(reinterpret_cast<float4*>(p),
           make_float4(__fdiv_rn(v[0], count), __fdiv_rn(v[1], count), __fdiv_rn(v[2], count),
                       __fdiv_rn(v[3], count)));
}
template <>
__device__ __forceinline__ void store_stream<1>(float* p, const float (&v)[1], float count) {
    __stcs(p, __fdiv_rn(v[0], count));
}

// Replica stores (possibly to a peer GPU): default cache policy, write-back at L2 / posted over
// NVLink.
template <int VW>
__device__ __forceinline__ void store_plain(float* p, const float (&v)[VW], float count);
template <>
__device__ __forceinline__ void store_plain<4>(float* p, const float (&v)[4], float count) {
    *reinterpret_cast<float4*>(p) =
        make_float4(__fdiv_rn(v[0], count), __fdiv_rn(v[1], count), __fdiv_rn(v[2], count),
                    __fdiv_rn(v[3], count));
}
template <>
__device__ __forceinline__ void store_plain<1>(float* p, const float (&v)[1], float count) {
    *p = __fdiv_rn(v[0], count);
}

// ---- the warp's view of its contiguous colidx span --------------------------------------
// Positions are 32-bit offsets from the 32-aligned arc position at or below the span start (a
// span is at most 31 rows; gr_csr_create rejects graphs whose 32-row windows exceed 2^31 arcs).
struct ArcStream {
    const int32_t* __restrict__ span;  // colidx + aligned span start
    uint32_t limit;                    // span length: offsets >= limit are never dereferenced
    uint32_t base;                     // offset of the chunk held in `cur` (multiple of 32)
    int32_t cur, nxt;

    __device__ __forceinline__ int32_t fetch(uint32_t p) const {
        return p < limit ? __ldcs(span + p) : 0;
    }
    __device__ __forceinline__ void open(uint32_t pos, int lane) {
        base = pos & ~31u;
        cur = fetch(base + lane);
        nxt = fetch(base + 32 + lane);
    }
    // make `pos` fall inside the current chunk (warp-uniform control flow)
    __device__ __forceinline__ void seek(uint32_t pos, int lane) {
        if (pos < base + 32) return;
        if (pos < base + 64) {
            cur = nxt;
            base += 32;
            nxt = fetch(base + 32 + lane);
        } else {
            open(pos, lane);
        }
    }
};

// Sum of X[colidx[k], col..col+VW) over span offsets k in [beg, end); result replicated in
// every lane group.  One gather per lane in flight; 32/LPR lane groups -> that many
// independent fp32 accumulators per column.
template <int LPR, int VW, int U>
__device__ __forceinline__ void reduce_arcs(ArcStream& s, uint32_t beg, uint32_t end,
                                            const float* __restrict__ xcol, int64_t ldx,
                                            bool col_ok, int lane, uint64_t pol_hot,
                                            float (&total)[VW]) {
    constexpr int G = 32 / LPR;
    const int grp = lane / LPR;
    float acc[VW];
#pragma unroll
    for (int c = 0; c < VW; ++c) acc[c] = 0.f;

    uint32_t k = beg;
    while (k < end) {
        s.seek(k, lane);
        const int off = (int)(k - s.base);
        const int cnt = (int)min(end - k, (uint32_t)(32 - off));
        if constexpr (U == 1) {
#pragma unroll 4
            for (int t = 0; t < cnt; t += G) {   // warp-uniform trip count (shuffles inside)
                const int kk = t + grp;
                const int32_t tagged = __shfl_sync(kFull, s.cur, (off + kk) & 31);
                float v[VW];
                load_row<VW>(v, xcol + (int64_t)(tagged & 0x7fffffff) * ldx, col_ok && kk < cnt,
                             tagged < 0, pol_hot);
#pragma unroll
                for (int c = 0; c < VW; ++c) acc[c] += v[c];
            }
        } else {
            // two gathers per lane in flight (narrow rows: the same bytes in flight per SM as a
            // 256-byte row with one); the additions keep the order of the U == 1 loop
#pragma unroll 2
            for (int t = 0; t < cnt; t += 2 * G) {
                const int k0 = t + grp, k1 = t + G + grp;
                const int32_t tag0 = __shfl_sync(kFull, s.cur, (off + k0) & 31);
                const int32_t tag1 = __shfl_sync(kFull, s.cur, (off + k1) & 31);
                float v0[VW], v1[VW];
                load_row<VW>(v0, xcol + (int64_t)(tag0 & 0x7fffffff) * ldx, col_ok && k0 < cnt,
                             tag0 < 0, pol_hot);
                load_row<VW>(v1, xcol + (int64_t)(tag1 & 0x7fffffff) * ldx, col_ok && k1 < cnt,
                             tag1 < 0, pol_hot);
#pragma unroll
                for (int c = 0; c < VW; ++c) acc[c] += v0[c];
#pragma unroll
                for (int c = 0; c < VW; ++c) acc[c] += v1[c];
            }
        }
        k += cnt;
    }
#pragma unroll
    for (int c = 0; c < VW; ++c) {
        float t = acc[c];
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) t += __shfl_xor_sync(kFull, t, o);
        total[c] = t;
    }
}

// ---- the level kernel ----------------------------------------------------------------------
// CTAs [0, n_seg_blocks): one warp per hub segment (kHubSegment arcs of a long row -> fp32
// partial); they lead the grid so the long warps start first and overlap the ordinary rows.
// Remaining CTAs: rows_per_warp consecutive ordinary rows per warp.
template <int LPR, int VW, bool BCAST, int U>
__device__ __forceinline__ void gather_body(const RefexArgs& a, const Replicas* rep) {
    constexpr int G = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR;
    const int grp = lane / LPR;
    const int col = ((int)blockIdx.y * LPR + sub) * VW;
    const bool col_ok = col < a.d;
    const float* xcol = a.X + col;
    const uint64_t pol_hot = l2_policy_evict_last();
    ArcStream s;

    if ((int64_t)blockIdx.x < a.n_seg_blocks) {
        const int64_t seg = a.seg_lo + (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
        if (seg >= a.seg_hi) return;
        const int64_t beg = __ldg(a.seg_begin + seg), end = __ldg(a.seg_end + seg);
        const int64_t base = beg & ~int64_t(31);
        s.span = a.colidx + base;
        s.limit = (uint32_t)(end - base);
        s.open((uint32_t)(beg - base), lane);
        float tot[VW];
        reduce_arcs<LPR, VW, U>(s, (uint32_t)(beg - base), s.limit, xcol, a.ldx, col_ok, lane,
                             pol_hot, tot);
        if (col_ok && grp == 0) store_sum<VW>(a.partial + seg * a.d + col, tot);
        return;
    }

    const int64_t first = a.row_lo + (((int64_t)blockIdx.x - a.n_seg_blocks) * kWarps +
                                      (threadIdx.x >> 5)) * a.rows_per_warp;
    if (first >= a.row_hi) return;
    const int nrows = (int)min((int64_t)a.rows_per_warp, a.row_hi - first);
    const int64_t rp64 = lane <= nrows ? __ldg(a.rowptr + first + lane) : 0;
    // offsets are taken from the 32-aligned position at or below the span start, so chunk
    // boundaries (and with them the summation order) depend on absolute arc positions only,
    // not on how rows are grouped into warps or on the row range of the call
    const int64_t span_beg = __shfl_sync(kFull, rp64, 0) & ~int64_t(31);
    const uint32_t rp = (uint32_t)(rp64 - span_beg);   // span-relative row boundaries
    s.span = a.colidx + span_beg;
    s.limit = __shfl_sync(kFull, rp, nrows);
    s.open(__shfl_sync(kFull, rp, 0), lane);

    for (int r = 0; r < nrows; ++r) {
        const uint32_t beg = __shfl_sync(kFull, rp, r);
        const uint32_t end = __shfl_sync(kFull, rp, r + 1);
        const uint32_t deg = end - beg;
        if (deg > a.hub_threshold) continue;  // segment warps + hub_fixup_kernel
        float tot[VW];
        reduce_arcs<LPR, VW, U>(s, beg, end, xcol, a.ldx, col_ok, lane, pol_hot, tot);
        if (!col_ok) continue;
        const int64_t o = (first + r) * a.ldo + col;
        // lane group 0 writes the sum block, group 1 (or the same lanes when LPR == 32) the mean
        if (a.out_sum && grp == 0) store_sum<VW>(a.out_sum + o, tot);
        if constexpr (BCAST) {
            // lane group g writes replicas g, g + G, ...: remote stores are posted writes, the
            // groups only split the issue slots
            const float cnt = (float)max(deg, 1u);
            for (int p = grp; p < rep->n_rep; p += G) store_plain<VW>(rep->mean[p] + o, tot, cnt);
        } else {
            if (a.out_mean && grp == (G >= 2 ? 1 : 0))
                store_stream<VW>(a.out_mean + o, tot, (float)max(deg, 1u));
        }
    }
}

template <int LPR, int VW>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks)
refex_gather_kernel(const RefexArgs a) {
    gather_body<LPR, VW, false, 1>(a, nullptr);
}

// Same gather, mean rows broadcast to every replica (own + peers over NVLink).
template <int LPR, int VW>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks)
refex_gather_bcast_kernel(const RefexArgs a, const Replicas rep) {
    gather_body<LPR, VW, true, 1>(a, &rep);
}

// Narrow rows (at most 32 fp32 columns, e.g. a column-group shard): two gathers per lane in
// flight, 40 registers, 48 resident warps per SM.
constexpr int kMinBlocksU2 = 6;
template <int LPR, int VW>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocksU2)
refex_gather_u2_kernel(const RefexArgs a) {
    gather_body<LPR, VW, false, 2>(a, nullptr);
}
template <int LPR, int VW>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocksU2)
refex_gather_bcast_u2_kernel(const RefexArgs a, const Replicas rep) {
    gather_body<LPR, VW, true, 2>(a, &rep);
}

// One warp per hub row: add the row's segment partials in fp64, in segment order.
__global__ void __launch_bounds__(kWarps * 32)
hub_fixup_kernel(const int64_t* __restrict__ hub_row, const int64_t* __restrict__ hub_seg_first,
                 const int64_t* __restrict__ rowptr, int64_t hub_lo, int64_t hub_hi,
                 const float* __restrict__ partial, int32_t d, float* __restrict__ out_sum,
                 float* __restrict__ out_mean, int64_t ldo, const Replicas rep) {
    const int lane = threadIdx.x & 31;
    const int64_t h = hub_lo + (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (h >= hub_hi) return;
    const int64_t row = hub_row[h];
    const int64_t s0 = hub_seg_first[h], s1 = hub_seg_first[h + 1];
    const float deg = (float)(rowptr[row + 1] - rowptr[row]);
    for (int c = lane; c < d; c += 32) {
        double acc = 0.0;
        for (int64_t s = s0; s < s1; ++s) acc += (double)partial[s * d + c];
        // same rule as the ordinary rows: mean = fp32 sum / count, correctly rounded
        const float sum = (float)acc, mean = __fdiv_rn(sum, deg);
        if (out_sum) out_sum[row * ldo + c] = sum;
        if (out_mean) out_mean[row * ldo + c] = mean;
        for (int p = 0; p < rep.n_rep; ++p) rep.mean[p][row * ldo + c] = mean;
    }
}

// ---- launch plumbing --------------------------------------------------------------------
int env_int(const char* name, int dflt, int lo, int hi);

template <int LPR, int VW>
cudaError_t launch_gather(const RefexArgs& a, const Replicas* rep, dim3 row_grid, dim3 seg_grid,
                          cudaStream_t st) {
    dim3 grid(row_grid.x + seg_grid.x, row_grid.y, 1);
    // rows of at most 128 bytes: two gathers per lane (GR_REFEX_U=1 / 2 overrides)
    bool u2 = false;
    if constexpr (VW == 4 && (LPR == 4 || LPR == 8))
        u2 = env_int("GR_REFEX_U", kNarrowRowGathers, 1, 2) == 2;
    if constexpr (VW == 4 && (LPR == 4 || LPR == 8)) {
        if (u2) {
            if (rep)
                refex_gather_bcast_u2_kernel<LPR, VW><<<grid, kWarps * 32, 0, st>>>(a, *rep);
            else
                refex_gather_u2_kernel<LPR, VW><<<grid, kWarps * 32, 0, st>>>(a);
            count_launch();
            return cudaGetLastError();
        }
    }
    if (rep)
        refex_gather_bcast_kernel<LPR, VW><<<grid, kWarps * 32, 0, st>>>(a, *rep);
    else
        refex_gather_kernel<LPR, VW><<<grid, kWarps * 32, 0, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

template <int VW>
cudaError_t dispatch_lpr(int lpr, const RefexArgs& a, const Replicas* rep, dim3 rg, dim3 sg,
                         cudaStream_t st) {
    switch (lpr) {
        case 1: return launch_gather<1, VW>(a, rep, rg, sg, st);
        case 2: return launch_gather<2, VW>(a, rep, rg, sg, st);
        case 4: return launch_gather<4, VW>(a, rep, rg, sg, st);
        case 8: return launch_gather<8, VW>(a, rep, rg, sg, st);
        case 16: return launch_gather<16, VW>(a, rep, rg, sg, st);
        default: return launch_gather<32, VW>(a, rep, rg, sg, st);
    }
}

int env_int(const char* name, int dflt, int lo, int hi) {
    const char* s = getenv(name);
    if (!s || !*s) return dflt;
    const int v = atoi(s);
    return v < lo ? lo : (v > hi ? hi : v);
}

int ensure_floats(float** buf, size_t* have, size_t want) {
    if (*have >= want) return GR_OK;
    if (*buf) {
        GR_CUDA_TRY(cudaFree(*buf));
        *buf = nullptr;
        *have = 0;
    }
    GR_CUDA_TRY(cudaMalloc(buf, want * sizeof(float)));
    *have = want;
    return GR_OK;
}

}  // namespace

namespace {

// One level over rows [row_lo, row_hi) of the handle.  rep == nullptr: plain outputs; otherwise
// the mean rows go to every replica in *rep (out_mean is ignored) -- the fused exchange.
int aggregate_impl(gr_csr_t* g, const float* X, int64_t ldx, int32_t d, int64_t row_lo,
                   int64_t row_hi, float* out_sum, float* out_mean, int64_t ldo,
                   const Replicas* rep, void* stream) {
    GR_REQUIRE(g != nullptr, "gr_refex_aggregate_f32: handle is NULL");
    GR_REQUIRE(d >= 1, "gr_refex_aggregate_f32: d = %d, need d >= 1", d);
    GR_REQUIRE(X != nullptr, "gr_refex_aggregate_f32: X is NULL");
    GR_REQUIRE(out_sum != nullptr || out_mean != nullptr || rep != nullptr,
               "gr_refex_aggregate_f32: both outputs are NULL");
    GR_REQUIRE(ldx >= d && ldo >= d, "gr_refex_aggregate_f32: ldx = %lld / ldo = %lld < d = %d",
               (long long)ldx, (long long)ldo, d);
    GR_REQUIRE(0 <= row_lo && row_lo <= row_hi && row_hi <= g->n_rows,
               "gr_refex_aggregate_f32: row range [%lld, %lld) outside [0, %lld)",
               (long long)row_lo, (long long)row_hi, (long long)g->n_rows);
    if (row_lo == row_hi) return GR_OK;

    DeviceGuard guard(g->device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "cannot select device %d", g->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    // hub rows / segments intersecting the row range
    const auto hb = std::lower_bound(g->h_hub_row.begin(), g->h_hub_row.end(), row_lo);
    const auto he = std::lower_bound(g->h_hub_row.begin(), g->h_hub_row.end(), row_hi);
    const int64_t hub_lo = hb - g->h_hub_row.begin(), hub_hi = he - g->h_hub_row.begin();
    const int64_t seg_lo = g->h_hub_seg_first[(size_t)hub_lo];
    const int64_t seg_hi = g->h_hub_seg_first[(size_t)hub_hi];
    if (seg_hi > seg_lo)
        if (int rc = ensure_floats(&g->d_partial, &g->partial_floats,
                                   (size_t)g->n_segments * (size_t)d))
            return rc;

    RefexArgs a;
    a.rowptr = g->rowptr;
    a.colidx = g->d_colidx_tagged ? g->d_colidx_tagged : g->colidx;
    a.X = X;
    a.ldx = ldx;
    a.d = d;
    a.row_lo = row_lo;
    a.row_hi = row_hi;
    a.out_sum = out_sum;
    a.out_mean = out_mean;
    a.ldo = ldo;
    a.seg_begin = g->d_seg_begin;
    a.seg_end = g->d_seg_end;
    a.seg_lo = seg_lo;
    a.seg_hi = seg_hi;
    a.partial = g->d_partial;
    a.n_seg_blocks = ceil_div<int64_t>(seg_hi - seg_lo, kWarps);
    // 16 consecutive rows per warp is the tuned point for ~40 arcs per row (C3).  A shard made of
    // the heavy rows of a power-law graph (the first node range) would serialise up to
    // 16 x hub_threshold arcs in one warp and its tail would outlast the rest of the launch
    // (measured: 7.7 ms for a quarter of C3 instead of 4.4), so long-row handles get fewer rows
    // per warp.  The grouping never changes the bits (chunking follows absolute arc positions).
    const int64_t avg_deg = g->nnz / std::max<int64_t>(1, g->n_rows);
    const int rpw_auto = avg_deg <= 64 ? 16 : (int)std::max<int64_t>(1, 1024 / avg_deg);
    a.rows_per_warp = env_int("GR_REFEX_ROWS_PER_WARP", rpw_auto, 1, 31);
    a.hub_threshold = (uint32_t)g->hub_threshold;

    bool vec4 = (d % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) && aligned16(X) &&
                (!out_sum || aligned16(out_sum)) && (!out_mean || aligned16(out_mean));
    if (rep)
        for (int p = 0; p < rep->n_rep; ++p) vec4 = vec4 && aligned16(rep->mean[p]);
    const int vw = vec4 ? 4 : 1;
    const int units = ceil_div<int>(d, vw);  // lanes needed to cover one feature row
    int lpr = vec4 ? 1 : 4;
    while (lpr < units && lpr < 32) lpr <<= 1;
    const int col_tiles = ceil_div<int>(units, lpr);

    const int64_t row_blocks =
        ceil_div<int64_t>(row_hi - row_lo, (int64_t)kWarps * a.rows_per_warp);
    GR_REQUIRE(row_blocks + a.n_seg_blocks < (int64_t)INT32_MAX && col_tiles <= 65535,
               "gr_refex_aggregate_f32: grid too large");
    dim3 row_grid((unsigned)row_blocks, (unsigned)col_tiles, 1);
    dim3 seg_grid((unsigned)a.n_seg_blocks, (unsigned)col_tiles, 1);

    cudaError_t e = vec4 ? dispatch_lpr<4>(lpr, a, rep, row_grid, seg_grid, st)
                         : dispatch_lpr<1>(lpr, a, rep, row_grid, seg_grid, st);
    if (e != cudaSuccess)
        return fail(GR_ERR_CUDA, "refex_gather_kernel launch failed: %s", cudaGetErrorString(e));

    if (hub_hi > hub_lo) {
        hub_fixup_kernel<<<(unsigned)ceil_div<int64_t>(hub_hi - hub_lo, kWarps), kWarps * 32, 0,
                           st>>>(g->d_hub_row, g->d_hub_seg_first, g->rowptr, hub_lo, hub_hi,
                                 g->d_partial, d, out_sum, rep ? nullptr : out_mean, ldo,
                                 rep ? *rep : Replicas{{}, 0});
        count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess)
            return fail(GR_ERR_CUDA, "hub_fixup_kernel launch failed: %s", cudaGetErrorString(e));
    }
    return GR_OK;
}

}  // namespace

extern "C" int gr_refex_aggregate_f32(gr_csr_t* g, const float* X, int64_t ldx, int32_t d,
                                      int64_t row_lo, int64_t row_hi, float* out_sum,
                                      float* out_mean, int64_t ldo, void* stream) {
    return aggregate_impl(g, X, ldx, d, row_lo, row_hi, out_sum, out_mean, ldo, nullptr, stream);
}

extern "C" int gr_refex_aggregate_bcast_f32(gr_csr_t* g, const float* X, int64_t ldx, int32_t d,
                                            int64_t row_lo, int64_t row_hi, float* out_sum,
                                            float* const* mean_replicas, int32_t n_replicas,
                                            int64_t ldo, void* stream) {
    GR_REQUIRE(mean_replicas != nullptr && n_replicas >= 1 && n_replicas <= kMaxReplicas,
               "gr_refex_aggregate_bcast_f32: need 1..%d replica pointers, got %d", kMaxReplicas,
               n_replicas);
    Replicas rep;
    rep.n_rep = n_replicas;
    for (int p = 0; p < kMaxReplicas; ++p) {
        rep.mean[p] = p < n_replicas ? mean_replicas[p] : nullptr;
        GR_REQUIRE(p >= n_replicas || rep.mean[p] != nullptr,
                   "gr_refex_aggregate_bcast_f32: replica %d is NULL", p);
        GR_REQUIRE(p >= n_replicas || rep.mean[p] != X,
                   "gr_refex_aggregate_bcast_f32: replica %d aliases X", p);
    }
    return aggregate_impl(g, X, ldx, d, row_lo, row_hi, out_sum, nullptr, ldo, &rep, stream);
}

extern "C" int gr_refex_levels_host_f32(gr_csr_t* g, const float* X_host, int64_t ldx, int32_t d,
                                        int32_t levels, int32_t recurse_on, float* out_host,
                                        void* stream) {
    GR_REQUIRE(g != nullptr, "gr_refex_levels_host_f32: handle is NULL");
    GR_REQUIRE(X_host != nullptr && out_host != nullptr, "gr_refex_levels_host_f32: NULL buffer");
    GR_REQUIRE(d >= 1 && ldx >= d && levels >= 1, "gr_refex_levels_host_f32: bad d/ldx/levels");
    GR_REQUIRE(recurse_on == 0 || recurse_on == 1, "recurse_on must be 0 (sum) or 1 (mean)");
    GR_REQUIRE(levels == 1 || g->n_rows == g->n_cols,
               "gr_refex_levels_host_f32: levels > 1 needs an unsharded graph "
               "(n_rows = %lld, n_cols = %lld)", (long long)g->n_rows, (long long)g->n_cols);

    DeviceGuard guard(g->device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "cannot select device %d", g->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    const size_t out_floats = (size_t)g->n_rows * 2 * (size_t)d;
    if (int rc = ensure_floats(&g->d_stage_x, &g->stage_x_floats, (size_t)g->n_cols * d)) return rc;
    if (g->stage_out_floats < out_floats) {
        for (int i = 0; i < 2; ++i) {
            size_t have = g->stage_out_floats;
            if (int rc = ensure_floats(&g->d_stage_out[i], &have, out_floats)) return rc;
        }
        g->stage_out_floats = out_floats;
    }
    if (!g->copy_stream) {
        GR_CUDA_TRY(cudaStreamCreateWithFlags(&g->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            GR_CUDA_TRY(cudaEventCreateWithFlags(&g->ev_level[i], cudaEventDisableTiming));
            GR_CUDA_TRY(cudaEventCreateWithFlags(&g->ev_copied[i], cudaEventDisableTiming));
        }
    }

    GR_CUDA_TRY(cudaMemcpy2DAsync(g->d_stage_x, (size_t)d * sizeof(float), X_host,
                                  (size_t)ldx * sizeof(float), (size_t)d * sizeof(float),
                                  (size_t)g->n_cols, cudaMemcpyHostToDevice, st));
    for (int l = 0; l < levels; ++l) {
        const int slot = l & 1;
        const float* in = l == 0 ? g->d_stage_x
                                 : g->d_stage_out[slot ^ 1] + (size_t)recurse_on * d;
        const int64_t ld_in = l == 0 ? d : 2 * (int64_t)d;
        float* out = g->d_stage_out[slot];
        // level l-2 used this slot: its device->host copy must have drained
        if (l >= 2) GR_CUDA_TRY(cudaStreamWaitEvent(st, g->ev_copied[slot], 0));
        if (int rc = gr_refex_aggregate_f32(g, in, ld_in, d, 0, g->n_rows, out, out + d,
                                            2 * (int64_t)d, st))
            return rc;
        GR_CUDA_TRY(cudaEventRecord(g->ev_level[slot], st));
        GR_CUDA_TRY(cudaStreamWaitEvent(g->copy_stream, g->ev_level[slot], 0));
        GR_CUDA_TRY(cudaMemcpyAsync(out_host + (size_t)l * out_floats, out,
                                    out_floats * sizeof(float), cudaMemcpyDeviceToHost,
                                    g->copy_stream));
        GR_CUDA_TRY(cudaEventRecord(g->ev_copied[slot], g->copy_stream));
    }
    GR_CUDA_TRY(cudaStreamSynchronize(g->copy_stream));
    GR_CUDA_TRY(cudaStreamSynchronize(st));
    return GR_OK;
}

// gr_peer_barrier lives in peer.cu
extern "C" int gr_peer_barrier(void* const* flag_arrays, int32_t n_ranks, int32_t rank,
                               int64_t epoch, double timeout_s, void* stream);

// Host-buffer entry point of a node-range shard (one call per rank, all ranks of the exchange
// group call it together): H2D of the whole level-0 input into the own replica, `levels` fused
// gather + broadcast levels separated by flag barriers, and a D2H copy of the OWN rows of every
// level on a second stream, overlapped with the next level.
extern "C" int gr_refex_levels_host_sharded_f32(
    gr_csr_t* g, const float* X_host, int64_t ldx, int32_t d, int32_t levels, int64_t row_offset,
    float* const* replicas_even, float* const* replicas_odd, void* const* flag_arrays,
    int32_t n_ranks, int32_t rank, int64_t* epoch_inout, float* out_host, void* stream) {
    GR_REQUIRE(g != nullptr, "gr_refex_levels_host_sharded_f32: handle is NULL");
    GR_REQUIRE(X_host && out_host && replicas_even && replicas_odd && flag_arrays && epoch_inout,
               "gr_refex_levels_host_sharded_f32: NULL argument");
    GR_REQUIRE(d >= 1 && ldx >= d && levels >= 1, "gr_refex_levels_host_sharded_f32: bad d/ldx/levels");
    GR_REQUIRE(n_ranks >= 1 && n_ranks <= kMaxReplicas && rank >= 0 && rank < n_ranks,
               "gr_refex_levels_host_sharded_f32: rank %d of %d", rank, n_ranks);
    GR_REQUIRE(row_offset >= 0 && row_offset + g->n_rows <= g->n_cols,
               "gr_refex_levels_host_sharded_f32: rows [%lld, %lld) outside [0, %lld)",
               (long long)row_offset, (long long)(row_offset + g->n_rows), (long long)g->n_cols);
    DeviceGuard guard(g->device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "cannot select device %d", g->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    // two local sum buffers (level l's D2H drains while level l + 1 runs)
    const size_t sum_floats = (size_t)std::max<int64_t>(g->n_rows, 1) * (size_t)d;
    if (g->stage_out_floats < sum_floats) {
        for (int i = 0; i < 2; ++i) {
            size_t have = g->stage_out_floats;
            if (int rc = ensure_floats(&g->d_stage_out[i], &have, sum_floats)) return rc;
        }
        g->stage_out_floats = sum_floats;
    }
    if (!g->copy_stream) {
        GR_CUDA_TRY(cudaStreamCreateWithFlags(&g->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            GR_CUDA_TRY(cudaEventCreateWithFlags(&g->ev_level[i], cudaEventDisableTiming));
            GR_CUDA_TRY(cudaEventCreateWithFlags(&g->ev_copied[i], cudaEventDisableTiming));
        }
    }

    float* const* reps[2] = {replicas_even, replicas_odd};
    const size_t row_bytes = (size_t)d * sizeof(float);
    // Level-0 input: every rank copies only ITS rows of X from the host (one PCIe link each) and
    // then pushes them into the peers' replicas over NVLink -- an all-gather of X0 that moves
    // n_cols * d * 4 / n_ranks bytes per PCIe link instead of all of it.
    if (g->n_rows > 0) {
        float* own = reps[0][rank] + (size_t)row_offset * d;
        GR_CUDA_TRY(cudaMemcpy2DAsync(own, row_bytes, X_host + (size_t)row_offset * ldx,
                                      (size_t)ldx * sizeof(float), row_bytes, (size_t)g->n_rows,
                                      cudaMemcpyHostToDevice, st));
        for (int p = 0; p < n_ranks; ++p)
            if (p != rank)
                GR_CUDA_TRY(cudaMemcpyAsync(reps[0][p] + (size_t)row_offset * d, own,
                                            (size_t)g->n_rows * row_bytes, cudaMemcpyDeviceToDevice,
                                            st));
    }
    if (n_ranks > 1) {
        *epoch_inout += 1;
        if (int rc = gr_peer_barrier(flag_arrays, n_ranks, rank, *epoch_inout, 0.0, st)) return rc;
    }
    const size_t level_floats = (size_t)g->n_rows * 2 * (size_t)d;
    for (int l = 0; l < levels; ++l) {
        const int slot = l & 1;
        const float* in = reps[slot][rank];
        float* const* outs = reps[slot ^ 1];
        float* sums = g->d_stage_out[slot];
        // level l - 2 used this sum buffer, and its D2H read the own rows of the replica this
        // level's kernel is about to overwrite
        if (l >= 2) GR_CUDA_TRY(cudaStreamWaitEvent(st, g->ev_copied[slot], 0));
        Replicas rep;
        rep.n_rep = n_ranks;
        for (int p = 0; p < kMaxReplicas; ++p)
            rep.mean[p] = p < n_ranks ? outs[p] + (size_t)row_offset * d : nullptr;
        if (int rc = aggregate_impl(g, in, d, d, 0, g->n_rows, sums, nullptr, d, &rep, st))
            return rc;
        GR_CUDA_TRY(cudaEventRecord(g->ev_level[slot], st));
        if (n_ranks > 1) {
            *epoch_inout += 1;
            if (int rc = gr_peer_barrier(flag_arrays, n_ranks, rank, *epoch_inout, 0.0, st))
                return rc;
        }
        if (g->n_rows > 0) {
            GR_CUDA_TRY(cudaStreamWaitEvent(g->copy_stream, g->ev_level[slot], 0));
            // two contiguous copies (a pitched copy of 128-byte rows runs at a fraction of the
            // PCIe rate): the level's sum rows, then its mean rows
            float* dst = out_host + (size_t)l * level_floats;
            const size_t block_bytes = (size_t)g->n_rows * row_bytes;
            GR_CUDA_TRY(cudaMemcpyAsync(dst, sums, block_bytes, cudaMemcpyDeviceToHost,
                                        g->copy_stream));
            GR_CUDA_TRY(cudaMemcpyAsync(dst + (size_t)g->n_rows * d,
                                        outs[rank] + (size_t)row_offset * d, block_bytes,
                                        cudaMemcpyDeviceToHost, g->copy_stream));
        }
        GR_CUDA_TRY(cudaEventRecord(g->ev_copied[slot], g->copy_stream));
    }
    GR_CUDA_TRY(cudaStreamSynchronize(g->copy_stream));
    GR_CUDA_TRY(cudaStreamSynchronize(st));
    return GR_OK;
}
