// RolX epilogue on the device ("next" #4): what RoleExtractor runs after the
// NMF for every cell of its (n_roles, n_bits) grid --
//
//   encode             graphrole/roles/factor.py:29-49: Lloyd-Max quantiser = 1-D
//                      KMeans(n_clusters=n_bins, random_state=1) over all entries of a factor
//                      (sklearn/cluster/_kmeans.py: k-means++ :180-278, Lloyd :620-758)
//   get_encoding_cost  graphrole/roles/description_length.py:32-41 (codebook size x entries)
//   get_error_cost     graphrole/roles/description_length.py:44-61 (generalised KL of V and G.F)
//   roles / role_percentage  graphrole/roles/extract.py:38-57 (row argmax / row normalisation)
//
// One-dimensional k-means has structure the generic algorithm does not use: after ONE sort of the
// values and ONE prefix sum, a Lloyd iteration is k binary searches (cluster = contiguous range of
// the sorted values) and k prefix-sum differences -- microseconds, independent of n -- so the whole
// Lloyd loop, its convergence tests and the empty-cluster relocation run inside a single one-CTA
// kernel launch, and the sort is shared by all the n_bins values the grid tries on one factor.
// sklearn's k-means++ seeding samples positions of the ORIGINAL order proportionally to the
// squared distance to the nearest chosen centre, which needs sums over that order: one streaming
// pass (8 bytes per entry) per centre for block sums of those distances.  Everything else about a
// new centre is local in the SORTED order: a candidate c between the existing centres c_L < c <
// c_R can only win points with c_L < x < c_R (for x <= c_L sklearn's rounded squared distance to
// c exceeds the one to c_L as soon as (c - c_L)^2 is above the rounding noise of the expansion
// x^2 - 2xc + c^2; closer candidates fall back to the whole range), so candidate potentials and
// the distance update touch only that range of the sorted values (through the sort
// permutation) -- about 2n/k entries per centre instead of n per candidate.  At 256 bins this
// is 40 x less traffic than round 2's first form (one scan, one fused candidate evaluation and
// one update pass over all entries per centre).
// The random stream is NumPy's RandomState(seed) (mt19937.h), so labels equal scikit-learn's
// whenever the data hold at least n_bins distinct values (otherwise scikit-learn's own result
// depends on np.argpartition's order of equal keys; here every distinct value becomes a level).

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <new>
#include <vector>

#include "common.cuh"
#include "mt19937.h"

using namespace gr;

namespace {

constexpr int kMaxBins = 1024;
constexpr int kMaxTrials = 8;          // 2 + int(ln(1024))
constexpr int kRedBlocks = 148 * 4;    // fixed grid of the deterministic two-stage reductions
constexpr int kRedThreads = 256;
constexpr int kPpBlock = 4096;         // k-means++: entries per block sum (original order)
constexpr int kPpStreamRounds = 20;    // centres 1 .. 19 are evaluated by streaming passes (measured: a range gather of
                                       // 4 - 7 candidates costs more than the stream while ranges hold > ~N/12 entries)

}  // namespace

struct gr_quantizer {
    int device = 0;
    int64_t capacity = 0;
    // bound matrix
    const void* src = nullptr;
    int is_f64 = 0;
    int64_t rows = 0, cols = 0, ld = 0, N = 0;
    double mean = 0.0, var = 0.0;
    // device buffers (capacity entries each)
    double* x = nullptr;        // centred values, original (row-major) order
    double* xs = nullptr;       // the same values sorted ascending
    double* P = nullptr;        // [N + 1] exclusive prefix sums of xs
    double* closest = nullptr;  // k-means++: squared distance to the nearest chosen centre
    int64_t* perm = nullptr;    // sorted position -> original position
    double* bsum = nullptr;     // k-means++: (cumulative) sums of `closest` per kPpBlock entries
    double* centres_sorted = nullptr;  // [kMaxBins] chosen centres, ascending
    // NumPy-ordered mean: the leaf table of its pairwise recursion for np_N entries (kept across
    // binds of equally sized matrices) and the leaf sums
    int64_t np_N = 0, np_leaves = 0;
    int64_t* np_offs = nullptr;
    int32_t* np_lens = nullptr;
    double* np_out = nullptr;
    std::vector<double> np_host;
    void* cub_temp = nullptr;
    size_t cub_temp_bytes = 0;
    double* partial = nullptr;  // [kRedBlocks * kMaxTrials]
    double* small = nullptr;    // device scratch: results of the reductions, candidates, ...
    int64_t* small_i = nullptr;
    double* lloyd_centers = nullptr;   // [kMaxBins]
    double* thresholds = nullptr;      // [kMaxBins] first sorted value of every used cluster
    double* levels = nullptr;          // [kMaxBins] centre (+ mean) of every used cluster
    int32_t* lloyd_info = nullptr;     // [4] n_iter, n_used, strict, relocations
    int64_t* lloyd_counts = nullptr;   // [kMaxBins]
};

namespace {

// ---- deterministic reductions ---------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double warp_part[kRedThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kRedThreads / 32; ++w) t += warp_part[w];
    return t;   // valid in thread 0
}

template <typename T>
__device__ __forceinline__ double load_value(const T* src, int64_t i, int64_t cols, int64_t ld) {
    return (double)src[(i / cols) * ld + (i % cols)];
}

// stage 1 of sum(v): per-block partials
template <typename T>
__global__ void __launch_bounds__(kRedThreads)
sum_kernel(const T* __restrict__ src, int64_t N, int64_t cols, int64_t ld,
           double* __restrict__ partial) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kRedThreads + threadIdx.x; i < N;
         i += (int64_t)gridDim.x * kRedThreads)
        acc += load_value(src, i, cols, ld);
    const double t = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// x = v - mean (sklearn centres the data, _kmeans.py:1486-1490) and sum x^2 for the variance
template <typename T>
__global__ void __launch_bounds__(kRedThreads)
center_kernel(const T* __restrict__ src, int64_t N, int64_t cols, int64_t ld, double mean,
              double* __restrict__ x, double* __restrict__ partial) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kRedThreads + threadIdx.x; i < N;
         i += (int64_t)gridDim.x * kRedThreads) {
        const double c = load_value(src, i, cols, ld) - mean;
        x[i] = c;
        acc += c * c;
    }
    const double t = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// ---- the data mean as NumPy computes it ------------------------------------------------------
// sklearn centres the data by X.mean(axis=0) (_kmeans.py:1486-1490), i.e. NumPy's pairwise sum
// (numpy/_core/src/umath/loops_utils.h.src: blocks of <= 128 values summed with eight running
// accumulators, halves split at a multiple of 8 and added pairwise).  On grid-valued data the
// k-means++ seeds are grid values, their midpoints are data values, and which centre such an
// exactly-equidistant point joins falls with the LAST BIT of that mean (measured: centres 1e-3
// apart after the first Lloyd step with a mean from an ordinary parallel reduction).  So the mean
// is summed in NumPy's order: one thread per leaf block here, the pairwise tree on the host.
template <typename T>
__global__ void __launch_bounds__(256)
numpy_leaf_sums_kernel(const T* __restrict__ src, int64_t cols, int64_t ld,
                       const int64_t* __restrict__ offs, const int32_t* __restrict__ lens,
                       int64_t n_leaves, double* __restrict__ out) {
    // eight lanes per leaf = NumPy's eight accumulators: lane j sums a[j], a[8 + j], a[16 + j], ...
    // in that order (64 contiguous bytes per step and leaf), then the fixed pairing
    // ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)) by xor shuffles (a + b == b + a bit for
    // bit, so every lane of a pair holds the same partial), then the tail on lane 0.
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t leaf = gid >> 3;
    const int j = (int)(gid & 7);
    if (leaf >= n_leaves) return;                 // whole 8-lane groups leave together
    const unsigned group_mask = 0xffu << ((threadIdx.x & 31) & ~7);
    const int64_t off = offs[leaf];
    const int n = lens[leaf];
    if (n < 8) {
        if (j == 0) {
            double r = 0.0;
            for (int i = 0; i < n; ++i) r = __dadd_rn(r, load_value(src, off + i, cols, ld));
            out[leaf] = r;
        }
        return;
    }
    double r = load_value(src, off + j, cols, ld);
    const int m = n - (n % 8);
    for (int i = 8; i < m; i += 8) r = __dadd_rn(r, load_value(src, off + i + j, cols, ld));
    r = __dadd_rn(r, __shfl_xor_sync(group_mask, r, 1, 8));
    r = __dadd_rn(r, __shfl_xor_sync(group_mask, r, 2, 8));
    r = __dadd_rn(r, __shfl_xor_sync(group_mask, r, 4, 8));
    if (j == 0) {
        for (int i = m; i < n; ++i) r = __dadd_rn(r, load_value(src, off + i, cols, ld));
        out[leaf] = r;
    }
}

// stage 2: out[t] = sum over blocks of partial[b * stride + t], fixed order
__global__ void finish_sum_kernel(const double* __restrict__ partial, int blocks, int stride,
                                  int count, double* __restrict__ out) {
    const int t = threadIdx.x;
    if (t >= count) return;
    double acc = 0.0;
    for (int b = 0; b < blocks; ++b) acc += partial[b * stride + t];
    out[t] = acc;
}

// ---- k-means++ ------------------------------------------------------------------------------
// sklearn's squared distance for one feature (pairwise.py:377-412 through _kmeans.py:240, :257):
// ((-2 (c x)) + c c) + x x, every operation rounded, clipped at 0.
__device__ __forceinline__ double sk_sqdist(double c, double cc, double x) {
    double d = -2.0 * __dmul_rn(c, x);
    d = __dadd_rn(d, cc);
    d = __dadd_rn(d, __dmul_rn(x, x));
    return fmax(d, 0.0);
}

__global__ void __launch_bounds__(kRedThreads)
pp_init_kernel(const double* __restrict__ x, int64_t N, double c, double* __restrict__ closest,
               double* __restrict__ partial) {
    const double cc = __dmul_rn(c, c);
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kRedThreads + threadIdx.x; i < N;
         i += (int64_t)gridDim.x * kRedThreads) {
        const double d = sk_sqdist(c, cc, x[i]);
        closest[i] = d;
        acc += d;
    }
    const double t = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void iota64_kernel(int64_t* p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// sums of `closest` over blocks of kPpBlock entries of the original order, fixed order
__global__ void __launch_bounds__(kRedThreads)
pp_block_sums_kernel(const double* __restrict__ closest, int64_t N, double* __restrict__ bsum) {
    const int64_t base = (int64_t)blockIdx.x * kPpBlock;
    double acc = 0.0;
    for (int j = threadIdx.x; j < kPpBlock; j += kRedThreads) {
        const int64_t i = base + j;
        if (i < N) acc += closest[i];
    }
    const double t = block_sum(acc);
    if (threadIdx.x == 0) bsum[blockIdx.x] = t;
}

// One CTA.  (1) bsum -> its inclusive cumulative sums, pot = their total (current_pot,
// _kmeans.py:246); (2) candidate_ids = clip(searchsorted(cumsum(closest), rand * pot), n - 1)
// (_kmeans.py:250-254): the block by binary search, the entry by a sequential sum inside the
// block; (3) per candidate the range [lo, hi) of SORTED positions its centre could win: the
// values strictly between the neighbouring chosen centres -- empty when the value is a centre
// already, everything when a neighbour is within rounding noise of it.
__global__ void __launch_bounds__(1024)
pp_scan_search_kernel(double* __restrict__ bsum, int nblk, const double* __restrict__ closest,
                      const double* __restrict__ x, const double* __restrict__ xs, int64_t N,
                      const double* __restrict__ u, int trials,
                      const double* __restrict__ centres_sorted, int n_centres,
                      int64_t* __restrict__ cand_idx, double* __restrict__ cand_x,
                      int64_t* __restrict__ cand_lo, int64_t* __restrict__ cand_hi,
                      double* __restrict__ pot_out) {
    __shared__ double slice_tot[1024];
    __shared__ double pot_s;
    int t = threadIdx.x;
    const int per = (nblk + 1023) / 1024;
    const int b0 = min(nblk, t * per), b1 = min(nblk, b0 + per);
    double run = 0.0;
    for (int b = b0; b < b1; ++b) { run += bsum[b]; bsum[b] = run; }
    slice_tot[t] = run;
    __syncthreads();
    if (t == 0) {
        double off = 0.0;
        for (int q = 0; q < 1024; ++q) { const double v = slice_tot[q]; slice_tot[q] = off; off += v; }
        pot_s = off;
        *pot_out = off;
    }
    __syncthreads();
    const double off = slice_tot[t];
    if (off != 0.0)
        for (int b = b0; b < b1; ++b) bsum[b] += off;
    __syncthreads();
    // one warp per candidate from here on
    const int w = t >> 5, lane = t & 31;
    if (w >= trials) return;
    const double pot = pot_s;
    const double v = __dmul_rn(u[w], pot);
    int lo_b = 0, hi_b = nblk;               // first block whose cumulative sum reaches v
    while (lo_b < hi_b) {
        const int mid = (lo_b + hi_b) >> 1;
        if (bsum[mid] < v) lo_b = mid + 1; else hi_b = mid;
    }
    int64_t idx = N - 1;
    if (lo_b < nblk) {
        // The entry inside the block, three levels of 32 so that no thread walks memory
        // serially (a single thread summing 4096 entries with an early exit cost ~1 ms per
        // centre): 32 segments of 128, the crossing segment as 32 groups of 4, then 4 entries.
        const double base = lo_b > 0 ? bsum[lo_b - 1] : 0.0;
        const int64_t i0 = (int64_t)lo_b * kPpBlock, i1 = min(N, i0 + kPpBlock);
        auto warp_prefix = [&](double s, double& excl) {       // fixed order, lane by lane
            double incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 0.0;
            return incl;
        };
        double s1 = 0.0;
        for (int j = 0; j < kPpBlock / 32; ++j) {
            const int64_t i = i0 + (int64_t)lane * (kPpBlock / 32) + j;
            if (i < i1) s1 += closest[i];
        }
        double excl1;
        const double incl1 = warp_prefix(s1, excl1);
        const unsigned m1 = __ballot_sync(0xffffffffu, base + incl1 >= v);
        idx = i1 - 1;
        if (m1 != 0) {
            const int l1 = __ffs(m1) - 1;
            const double acc1 = base + __shfl_sync(0xffffffffu, excl1, l1);
            const int64_t seg0 = i0 + (int64_t)l1 * (kPpBlock / 32);
            double e[4], s2 = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t i = seg0 + lane * 4 + j;
                e[j] = i < i1 ? closest[i] : 0.0;
                s2 += e[j];
            }
            double excl2;
            const double incl2 = warp_prefix(s2, excl2);
            const unsigned m2 = __ballot_sync(0xffffffffu, acc1 + incl2 >= v);
            idx = min(seg0 + kPpBlock / 32 - 1, i1 - 1);
            if (m2 != 0) {
                const int l2 = __ffs(m2) - 1;
                double acc2 = acc1 + __shfl_sync(0xffffffffu, excl2, l2);
                idx = min(seg0 + l2 * 4 + 3, i1 - 1);
                bool found = false;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc2 += __shfl_sync(0xffffffffu, e[j], l2);
                    if (!found && acc2 >= v) { idx = min(seg0 + l2 * 4 + j, i1 - 1); found = true; }
                }
            }
        }
    }
    if (lane != 0) return;
    t = w;
    const double c = x[idx];
    cand_idx[t] = idx;
    cand_x[t] = c;
    int pos = 0, hi_c = n_centres;           // first chosen centre >= c
    while (pos < hi_c) {
        const int mid = (pos + hi_c) >> 1;
        if (centres_sorted[mid] < c) pos = mid + 1; else hi_c = mid;
    }
    int64_t lo = 0, hi = 0;
    if (!(pos < n_centres && centres_sorted[pos] == c)) {
        const double m2 = fmax(xs[0] * xs[0], xs[N - 1] * xs[N - 1]);
        const double guard = 1e-14 * m2;
        const bool has_l = pos > 0, has_r = pos < n_centres;
        const double cl = has_l ? centres_sorted[pos - 1] : 0.0;
        const double cr = has_r ? centres_sorted[pos] : 0.0;
        const bool everything = (has_l && (c - cl) * (c - cl) <= guard) ||
                                (has_r && (cr - c) * (cr - c) <= guard);
        lo = 0;
        hi = N;
        if (!everything) {
            if (has_l) {                     // first sorted value > cl
                int64_t a = 0, b = N;
                while (a < b) { const int64_t m = a + (b - a) / 2; if (xs[m] <= cl) a = m + 1; else b = m; }
                lo = a;
            }
            if (has_r) {                     // first sorted value >= cr
                int64_t a = 0, b = N;
                while (a < b) { const int64_t m = a + (b - a) / 2; if (xs[m] < cr) a = m + 1; else b = m; }
                hi = a;
            }
        }
    }
    cand_lo[t] = lo;
    cand_hi[t] = hi;
}

// what candidate t would take off the potential: sum over its range of
// closest - min(closest, dist(c_t, x))  (candidates_pot = pot - that, _kmeans.py:257-263)
__global__ void __launch_bounds__(kRedThreads)
pp_eval_range_kernel(const double* __restrict__ xs, const int64_t* __restrict__ perm,
                     const double* __restrict__ closest, const double* __restrict__ cand_x,
                     const int64_t* __restrict__ cand_lo, const int64_t* __restrict__ cand_hi,
                     double* __restrict__ partial) {
    const int t = blockIdx.y;
    const double c = cand_x[t], cc = __dmul_rn(c, c);
    const int64_t hi = cand_hi[t];
    double acc = 0.0;
    for (int64_t p = cand_lo[t] + (int64_t)blockIdx.x * kRedThreads + threadIdx.x; p < hi;
         p += (int64_t)gridDim.x * kRedThreads) {
        const double cl = closest[perm[p]];
        acc += cl - fmin(cl, sk_sqdist(c, cc, xs[p]));
    }
    const double s = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x * kMaxTrials + t] = s;
}

// The same two steps as streaming passes over ALL entries in their original order: what the first
// few centres use, whose ranges are most of the data (a gather through the permutation moves a
// 32-byte sector per 8-byte value; a stream does not).
template <int TRIALS>
__global__ void __launch_bounds__(kRedThreads)
pp_eval_all_kernel(const double* __restrict__ x, const double* __restrict__ closest, int64_t N,
                   const double* __restrict__ cand_x, double* __restrict__ partial) {
    double c[TRIALS], cc[TRIALS], acc[TRIALS];
#pragma unroll
    for (int t = 0; t < TRIALS; ++t) {
        c[t] = cand_x[t];
        cc[t] = __dmul_rn(c[t], c[t]);
        acc[t] = 0.0;
    }
    for (int64_t i = (int64_t)blockIdx.x * kRedThreads + threadIdx.x; i < N;
         i += (int64_t)gridDim.x * kRedThreads) {
        const double xi = x[i], cl = closest[i];
#pragma unroll
        for (int t = 0; t < TRIALS; ++t) acc[t] += cl - fmin(cl, sk_sqdist(c[t], cc[t], xi));
    }
#pragma unroll
    for (int t = 0; t < TRIALS; ++t) {
        const double s = block_sum(acc[t]);
        if (threadIdx.x == 0) partial[blockIdx.x * kMaxTrials + t] = s;
    }
}

__global__ void __launch_bounds__(kRedThreads)
pp_update_all_kernel(const double* __restrict__ x, int64_t N, double c,
                     double* __restrict__ closest) {
    const double cc = __dmul_rn(c, c);
    for (int64_t i = (int64_t)blockIdx.x * kRedThreads + threadIdx.x; i < N;
         i += (int64_t)gridDim.x * kRedThreads)
        closest[i] = fmin(closest[i], sk_sqdist(c, cc, x[i]));
}

// closest = min(closest, dist(c, x)) over the chosen candidate's range  (_kmeans.py:266-270)
__global__ void __launch_bounds__(kRedThreads)
pp_update_range_kernel(const double* __restrict__ xs, const int64_t* __restrict__ perm, double c,
                       int64_t lo, int64_t hi, double* __restrict__ closest) {
    const double cc = __dmul_rn(c, c);
    for (int64_t p = lo + (int64_t)blockIdx.x * kRedThreads + threadIdx.x; p < hi;
         p += (int64_t)gridDim.x * kRedThreads) {
        const int64_t i = perm[p];
        closest[i] = fmin(closest[i], sk_sqdist(c, cc, xs[p]));
    }
}

// ---- the Lloyd loop: one CTA, everything in shared memory --------------------------------------
// E-step score of sklearn's chunked Lloyd iteration (_k_means_lloyd.pyx: |c|^2 - 2 x.c through a
// K = 1 GEMM): the smaller score wins, equal scores go to the smaller cluster index.
__device__ __forceinline__ double sk_score(double c, double x) {
    return __dadd_rn(__dmul_rn(c, c), -2.0 * __dmul_rn(x, c));
}

struct LloydShared {
    double c[kMaxBins];        // centres by cluster index
    double c_old[kMaxBins];
    double cs[kMaxBins];       // centres in sorted order
    double sum[kMaxBins];      // by sorted position (relocation edits it)
    int64_t start[kMaxBins + 1];      // E step: first sorted value of the range at a position
    int64_t start_prev[kMaxBins + 1]; // previous E step, for the "labels unchanged" test
    int64_t cnt_e[kMaxBins];   // E-step sizes by sorted position
    int64_t cnt[kMaxBins];     // sizes after relocation
    int ord[kMaxBins];         // sorted position -> cluster index
    int ord_prev[kMaxBins];
    int pos[kMaxBins];         // cluster index -> sorted position
    int strip_lo[kMaxBins], strip_hi[kMaxBins];   // points relocation took from a range's ends
    unsigned char active[kMaxBins];
    double red[32];
    int ired[32];
};

// sorted order of the centres by (value, cluster index): rank by counting, k <= 1024
__device__ void lloyd_sort_centres(LloydShared& s, int k) {
    const int j = threadIdx.x;
    if (j < k) {
        const double cj = s.c[j];
        int rank = 0;
        for (int i = 0; i < k; ++i) {
            const double ci = s.c[i];
            rank += (ci < cj || (ci == cj && i < j)) ? 1 : 0;
        }
        s.cs[rank] = cj;
        s.ord[rank] = j;
        s.pos[j] = rank;
    }
    __syncthreads();
    // equal centres: only the first (smallest cluster index) can win a point
    if (j < k) s.active[j] = (j == 0 || s.cs[j] != s.cs[j - 1]) ? 1 : 0;
    __syncthreads();
}

// E step on the sorted values: range of sorted position p = [start[p], start[p + 1]).
__device__ void lloyd_assign(LloydShared& s, int k, const double* __restrict__ xs, int64_t N) {
    const int p = threadIdx.x;
    if (p < k) {
        int64_t b = 0;
        if (s.active[p]) {
            int q = p - 1;                        // previous active centre
            while (q >= 0 && !s.active[q]) --q;
            if (q >= 0) {
                const double lo_c = s.cs[q], hi_c = s.cs[p];
                const int lo_i = s.ord[q], hi_i = s.ord[p];
                int64_t lo = 0, hi = N;           // first value that prefers the upper centre
                while (lo < hi) {
                    const int64_t mid = lo + (hi - lo) / 2;
                    const double xv = xs[mid];
                    const double su = sk_score(hi_c, xv), sl = sk_score(lo_c, xv);
                    const bool upper = su < sl || (su == sl && hi_i < lo_i);
                    if (upper) hi = mid; else lo = mid + 1;
                }
                b = lo;
            }
        }
        s.start[p] = b;
    }
    if (p == 0) s.start[k] = N;
    __syncthreads();
    if (p == 0) {
        // boundaries non-decreasing; a position whose centre duplicates an earlier one is empty
        int64_t run = 0;
        for (int q = 0; q < k; ++q) {
            if (s.active[q]) run = max(run, s.start[q]);
            s.start[q] = run;
        }
        for (int q = k - 1; q >= 0; --q)
            if (!s.active[q]) s.start[q] = s.start[q + 1];
    }
    __syncthreads();
    if (p < k) s.cnt_e[p] = s.start[p + 1] - s.start[p];
    __syncthreads();
}

// info: [0] n_iter  [1] number of used clusters  [2] strict convergence  [3] relocations
__global__ void __launch_bounds__(kMaxBins)
lloyd_kernel(const double* __restrict__ xs, const double* __restrict__ P, int64_t N, int k,
             double* __restrict__ centres, double tol, int max_iter, double mean,
             int32_t* __restrict__ info, int64_t* __restrict__ counts_out,
             double* __restrict__ thresholds, double* __restrict__ levels) {
    extern __shared__ __align__(16) unsigned char lloyd_smem[];
    LloydShared& s = *reinterpret_cast<LloydShared*>(lloyd_smem);
    const int p = threadIdx.x;
    const int n_warps = (int)(blockDim.x >> 5);
    if (p < k) {
        s.c[p] = centres[p];
        s.ord_prev[p] = -1;
    }
    if (p <= k) s.start_prev[p] = -1;
    __syncthreads();

    bool strict = false;
    int n_iter = 0, relocations = 0;
    for (int it = 0; it < max_iter; ++it) {
        n_iter = it + 1;
        lloyd_sort_centres(s, k);
        lloyd_assign(s, k, xs, N);
        if (p < k) {
            s.sum[p] = P[s.start[p + 1]] - P[s.start[p]];
            s.cnt[p] = s.cnt_e[p];
            s.strip_lo[p] = s.strip_hi[p] = 0;
            s.c_old[p] = s.c[p];
        }
        __syncthreads();

        // empty clusters take the points farthest from their E-step centres
        // (_k_means_common.pyx, _relocate_empty_clusters_dense); serial and rare
        if (p == 0) {
            for (int j = 0; j < k; ++j) {          // np.where(weight == 0)[0] order
                const int pj = s.pos[j];
                if (s.cnt_e[pj] != 0) continue;
                double best = -1.0;
                int bq = -1;
                bool at_lo = true;
                for (int q = 0; q < k; ++q) {
                    const int64_t left = s.cnt_e[q] - s.strip_lo[q] - s.strip_hi[q];
                    if (left <= 0) continue;
                    const double cq = s.c_old[s.ord[q]];
                    const double xl = xs[s.start[q] + s.strip_lo[q]];
                    const double xh = xs[s.start[q + 1] - 1 - s.strip_hi[q]];
                    const double dl = (xl - cq) * (xl - cq), dh = (xh - cq) * (xh - cq);
                    if (dl > best) { best = dl; bq = q; at_lo = true; }
                    if (dh > best) { best = dh; bq = q; at_lo = false; }
                }
                if (bq < 0) break;
                const double xv = at_lo ? xs[s.start[bq] + s.strip_lo[bq]]
                                        : xs[s.start[bq + 1] - 1 - s.strip_hi[bq]];
                if (at_lo) s.strip_lo[bq] += 1; else s.strip_hi[bq] += 1;
                s.sum[bq] -= xv;
                s.cnt[bq] -= 1;
                s.sum[pj] = xv;
                s.cnt[pj] = 1;
                ++relocations;
            }
        }
        __syncthreads();

        // M step (_average_centers: centre = sum * (1 / weight); a cluster left without weight
        // keeps its raw sum) and the centre shift
        double sh2 = 0.0;
        int same = 1;
        if (p < k) {
            const int j = s.ord[p];
            if (s.cnt[p] > 0) s.c[j] = s.sum[p] * (1.0 / (double)s.cnt[p]);
            else if (s.cnt_e[p] > 0) s.c[j] = s.sum[p];
            const double dlt = s.c[j] - s.c_old[j];
            sh2 = dlt * dlt;
            // labels unchanged <=> every non-empty range existed, with the same cluster index, in
            // the previous E step (both sets of ranges tile [0, N))
            if (s.cnt_e[p] > 0) {
                same = 0;
                for (int q = 0; q < k; ++q)
                    if (s.start_prev[q] == s.start[p] && s.start_prev[q + 1] == s.start[p + 1] &&
                        s.ord_prev[q] == j) { same = 1; break; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sh2 += __shfl_xor_sync(0xffffffffu, sh2, o);
        const unsigned warp_same = __all_sync(0xffffffffu, same != 0);
        if ((p & 31) == 0) {
            s.red[p >> 5] = sh2;
            s.ired[p >> 5] = warp_same ? 1 : 0;
        }
        __syncthreads();
        double shift_tot = 0.0;
        int all_same = 1;
        for (int w = 0; w < n_warps; ++w) {
            shift_tot += s.red[w];
            all_same &= s.ired[w];
        }
        __syncthreads();
        if (all_same) { strict = true; break; }        // _kmeans.py:723-728
        if (shift_tot <= tol) break;                   // :731-738
        if (p < k) s.ord_prev[p] = s.ord[p];
        if (p <= k) s.start_prev[p] = s.start[p];
        __syncthreads();
    }
    // labels that go with the final centres: the last E step under strict convergence, otherwise
    // one more E step (_kmeans.py:742-754)
    if (!strict) {
        lloyd_sort_centres(s, k);
        lloyd_assign(s, k, xs, N);
    }
    if (p < k) counts_out[s.ord[p]] = s.cnt_e[p];
    if (p == 0) {
        int m = 0;
        for (int q = 0; q < k; ++q) {
            if (s.cnt_e[q] <= 0) continue;
            thresholds[m] = xs[s.start[q]];
            levels[m] = s.c[s.ord[q]] + mean;          // best_centers += X_mean, :1546
            ++m;
        }
        info[0] = n_iter;
        info[1] = m;
        info[2] = strict ? 1 : 0;
        info[3] = relocations;
    }
    if (p < k) centres[p] = s.c[p] + mean;
}

// out[i] = level of the range that holds x[i] (graphrole/roles/factor.py:48)
template <typename T>
__global__ void __launch_bounds__(256)
quantize_kernel(const double* __restrict__ x, int64_t N, int64_t cols, int64_t ldo,
                const double* __restrict__ thresholds, const double* __restrict__ levels,
                int n_used, T* __restrict__ out) {
    __shared__ double th[kMaxBins], lv[kMaxBins];
    for (int i = threadIdx.x; i < n_used; i += blockDim.x) {
        th[i] = thresholds[i];
        lv[i] = levels[i];
    }
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double xv = x[i];
        int lo = 0, hi = n_used;                    // last m with th[m] <= xv
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (th[mid] <= xv) lo = mid; else hi = mid;
        }
        out[(i / cols) * ldo + (i % cols)] = (T)lv[lo];
    }
}

// fewer distinct values than bins: every distinct value is its own level (out = in)
template <typename T>
__global__ void __launch_bounds__(256)
copy_matrix_kernel(const T* __restrict__ src, int64_t N, int64_t cols, int64_t ld, int64_t ldo,
                   T* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N;
         i += (int64_t)gridDim.x * blockDim.x)
        out[(i / cols) * ldo + (i % cols)] = src[(i / cols) * ld + (i % cols)];
}

// number of distinct values of a sorted array
__global__ void __launch_bounds__(kRedThreads)
count_runs_kernel(const double* __restrict__ xs, int64_t N, double* __restrict__ partial) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kRedThreads + threadIdx.x; i < N;
         i += (int64_t)gridDim.x * kRedThreads)
        acc += (i == 0 || xs[i] != xs[i - 1]) ? 1.0 : 0.0;
    const double t = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// ---- description length --------------------------------------------------------------------
// error cost: sum over v != 0 of v log(v / v') - v + v' with v' = (G F)[i, j]
// (description_length.py:44-61), fp64 arithmetic whatever the storage type.  One CTA per row
// block; F (r x f) is read through L1/L2 (r f <= 32 x 1024 values).
template <typename T>
__global__ void __launch_bounds__(kRedThreads)
kl_cost_kernel(const T* __restrict__ V, int64_t n, int f, int64_t ldv, const T* __restrict__ G,
               int64_t ldg, const T* __restrict__ F, int64_t ldf, int r,
               double* __restrict__ partial) {
    __shared__ double g_row[64];
    double acc = 0.0;
    for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < r) g_row[threadIdx.x] = (double)G[i * ldg + threadIdx.x];
        __syncthreads();
        for (int j = threadIdx.x; j < f; j += kRedThreads) {
            const double v = (double)V[i * ldv + j];
            if (v != 0.0) {
                double a = 0.0;
                for (int q = 0; q < r; ++q) a = fma(g_row[q], (double)__ldg(F + q * ldf + j), a);
                acc += (v * log(v / a) - v) + a;
            }
        }
    }
    const double t = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// same cost from an explicit approximation matrix (get_error_cost(V, V_approx))
template <typename T>
__global__ void __launch_bounds__(kRedThreads)
kl_pair_kernel(const T* __restrict__ V, const T* __restrict__ A, int64_t n, int f, int64_t ldv,
               int64_t lda, double* __restrict__ partial) {
    double acc = 0.0;
    const int64_t total = n * f;
    for (int64_t e = (int64_t)blockIdx.x * kRedThreads + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * kRedThreads) {
        const int64_t i = e / f, j = e % f;
        const double v = (double)V[i * ldv + j];
        if (v != 0.0) {
            const double a = (double)A[i * lda + j];
            acc += (v * log(v / a) - v) + a;
        }
    }
    const double t = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// roles / role_percentage (roles/extract.py:38-57): first maximum of every row, row / row sum
template <typename T>
__global__ void __launch_bounds__(256)
roles_kernel(const T* __restrict__ W, int64_t n, int r, int64_t ldw, int32_t* __restrict__ argmax,
             T* __restrict__ pct, int64_t ldp) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const T* row = W + i * ldw;
        T best = row[0], tot = row[0];
        int arg = 0;
        for (int q = 1; q < r; ++q) {
            const T v = row[q];
            tot += v;
            if (v > best) { best = v; arg = q; }
        }
        if (argmax) argmax[i] = arg;
        if (pct)
            for (int q = 0; q < r; ++q) pct[i * ldp + q] = row[q] / tot;
    }
}

int reduce_blocks(int64_t N) {
    return (int)std::max<int64_t>(1, std::min<int64_t>(kRedBlocks, ceil_div<int64_t>(N, kRedThreads)));
}

// sum of `count` per-block partial vectors -> host
int finish_to_host(gr_quantizer* q, int blocks, int stride, int count, double* host,
                   cudaStream_t st) {
    finish_sum_kernel<<<1, 32, 0, st>>>(q->partial, blocks, stride, count, q->small);
    count_launch();
    GR_CUDA_TRY(cudaGetLastError());
    GR_CUDA_TRY(cudaMemcpyAsync(host, q->small, count * sizeof(double), cudaMemcpyDeviceToHost, st));
    GR_CUDA_TRY(cudaStreamSynchronize(st));
    return GR_OK;
}

// NumPy's pairwise sum (see numpy_leaf_sums_kernel): the leaves of its recursion ...
void numpy_leaves(int64_t off, int64_t n, std::vector<int64_t>& offs, std::vector<int32_t>& lens) {
    if (n <= 128) {
        offs.push_back(off);
        lens.push_back((int32_t)n);
        return;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    numpy_leaves(off, n2, offs, lens);
    numpy_leaves(off + n2, n - n2, offs, lens);
}
// ... and its additions over the leaf sums, in the same order
double numpy_combine(const double* leaf, int64_t& next, int64_t n) {
    if (n <= 128) return leaf[next++];
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    const double a = numpy_combine(leaf, next, n2);
    const double b = numpy_combine(leaf, next, n - n2);
    return a + b;
}
template <typename T>
int numpy_sum(gr_quantizer* q, const T* X, int64_t N, int64_t cols, int64_t ld, double* out,
              cudaStream_t st) {
    if (q->np_N != N) {
        std::vector<int64_t> offs;
        std::vector<int32_t> lens;
        offs.reserve((size_t)(N / 64 + 2));
        lens.reserve((size_t)(N / 64 + 2));
        numpy_leaves(0, N, offs, lens);
        cudaFree(q->np_offs);
        cudaFree(q->np_lens);
        cudaFree(q->np_out);
        q->np_offs = nullptr;
        q->np_lens = nullptr;
        q->np_out = nullptr;
        q->np_N = 0;
        const size_t nl = offs.size();
        GR_CUDA_TRY(cudaMalloc((void**)&q->np_offs, nl * sizeof(int64_t)));
        GR_CUDA_TRY(cudaMalloc((void**)&q->np_lens, nl * sizeof(int32_t)));
        GR_CUDA_TRY(cudaMalloc((void**)&q->np_out, nl * sizeof(double)));
        GR_CUDA_TRY(cudaMemcpyAsync(q->np_offs, offs.data(), nl * sizeof(int64_t),
                                    cudaMemcpyHostToDevice, st));
        GR_CUDA_TRY(cudaMemcpyAsync(q->np_lens, lens.data(), nl * sizeof(int32_t),
                                    cudaMemcpyHostToDevice, st));
        GR_CUDA_TRY(cudaStreamSynchronize(st));      // the host vectors go out of scope
        q->np_leaves = (int64_t)nl;
        q->np_host.resize(nl);
        q->np_N = N;
    }
    const int64_t n_leaves = q->np_leaves;
    numpy_leaf_sums_kernel<T><<<(unsigned)ceil_div<int64_t>(n_leaves * 8, 256), 256, 0, st>>>(
        X, cols, ld, q->np_offs, q->np_lens, n_leaves, q->np_out);
    count_launch();
    GR_CUDA_TRY(cudaGetLastError());
    GR_CUDA_TRY(cudaMemcpyAsync(q->np_host.data(), q->np_out, (size_t)n_leaves * sizeof(double),
                                cudaMemcpyDeviceToHost, st));
    GR_CUDA_TRY(cudaStreamSynchronize(st));
    int64_t next = 0;
    *out = numpy_combine(q->np_host.data(), next, N);
    return GR_OK;
}

template <typename T>
int bind_impl(gr_quantizer* q, const T* X, int64_t rows, int64_t cols, int64_t ld, void* stream) {
    GR_REQUIRE(q != nullptr, "gr_quantizer_bind: handle is NULL");
    GR_REQUIRE(X != nullptr && rows >= 1 && cols >= 1 && ld >= cols,
               "gr_quantizer_bind: bad matrix (rows %lld, cols %lld, ld %lld)", (long long)rows,
               (long long)cols, (long long)ld);
    const int64_t N = rows * cols;
    GR_REQUIRE(N <= q->capacity, "gr_quantizer_bind: %lld entries exceed the capacity %lld",
               (long long)N, (long long)q->capacity);
    DeviceGuard guard(q->device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "cannot select device %d", q->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = reduce_blocks(N);

    double h = 0.0;
    if (int rc = numpy_sum(q, X, N, cols, ld, &h, st)) return rc;
    q->mean = h / (double)N;
    center_kernel<T><<<blocks, kRedThreads, 0, st>>>(X, N, cols, ld, q->mean, q->x, q->partial);
    count_launch();
    if (int rc = finish_to_host(q, blocks, 1, 1, &h, st)) return rc;
    q->var = h / (double)N;

    // sorted values and the sort permutation (k-means++ reaches `closest` through it)
    int64_t* iota = reinterpret_cast<int64_t*>(q->closest);     // free until the next encode
    iota64_kernel<<<(unsigned)ceil_div<int64_t>(N, 256), 256, 0, st>>>(iota, N);
    size_t need = q->cub_temp_bytes;
    GR_CUDA_TRY(cub::DeviceRadixSort::SortPairs(q->cub_temp, need, q->x, q->xs, iota, q->perm, N,
                                                0, 64, st));
    GR_CUDA_TRY(cudaMemsetAsync(q->P, 0, sizeof(double), st));
    need = q->cub_temp_bytes;
    GR_CUDA_TRY(cub::DeviceScan::InclusiveSum(q->cub_temp, need, q->xs, q->P + 1, N, st));
    count_launch(6);   // cub: radix-sort passes + scan
    q->src = X;
    q->is_f64 = sizeof(T) == 8;
    q->rows = rows;
    q->cols = cols;
    q->ld = ld;
    q->N = N;
    return GR_OK;
}

// RandomState.choice(n, p = 1/n): cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, u, 'right').
// The sequential cumsum of a constant is emulated only when floor(u n) is not provably the answer
// (and n is small enough to afford it); beyond 2^24 entries floor(u n) is used as is.
int64_t choice_uniform(int64_t n, double u) {
    int64_t i0 = std::min<int64_t>((int64_t)(u * (double)n), n - 1);
    const double slack = ((double)n + 8.0) * 2.220446049250313e-16;
    const bool safe_hi = (double)(i0 + 1) / (double)n * (1.0 - slack) > u;
    const bool safe_lo = i0 == 0 || (double)i0 / (double)n * (1.0 + slack) <= u;
    if ((safe_hi && safe_lo) || n > (1ll << 24)) return i0;
    const double c = 1.0 / (double)n;
    volatile double tot = 0.0;
    for (int64_t i = 0; i < n; ++i) tot = tot + c;
    const double last = tot;
    volatile double run = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        run = run + c;
        if (run / last > u) return i;
    }
    return n - 1;
}

template <typename T>
int encode_impl(gr_quantizer* q, int32_t n_bins, uint32_t seed, int32_t max_iter, double tol,
                T* out, int64_t ldo, double* centers_out, int32_t* n_iter_out,
                int64_t* n_distinct_out, void* stream) {
    GR_REQUIRE(q != nullptr && q->src != nullptr, "gr_quantizer_encode: no matrix is bound");
    GR_REQUIRE(q->is_f64 == (sizeof(T) == 8), "gr_quantizer_encode: dtype differs from bind");
    GR_REQUIRE(out != nullptr && ldo >= q->cols, "gr_quantizer_encode: bad output");
    GR_REQUIRE(n_bins >= 1 && n_bins <= kMaxBins, "gr_quantizer_encode: n_bins = %d outside [1, %d]",
               n_bins, kMaxBins);
    // the message sklearn raises (KMeans._check_params_vs_input); callers of the reference rely
    // on this ValueError to skip grid cells (roles/extract.py:127-129)
    GR_REQUIRE(q->N >= n_bins, "n_samples=%lld should be >= n_clusters=%d.", (long long)q->N, n_bins);
    DeviceGuard guard(q->device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "cannot select device %d", q->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t N = q->N;
    const int blocks = reduce_blocks(N);
    const int k = n_bins;
    const int trials = 2 + (int)std::log((double)k);

    // ---- k-means++ seeding (_kmeans.py:180-278) ------------------------------------------------
    NumpyRandomState rs(seed);
    std::vector<double> centres((size_t)k);
    double h[kMaxTrials];
    double* d_rand = q->small + 16;
    double* d_cand_x = q->small + 32;
    int64_t* d_cand_i = q->small_i;
    const int64_t first = choice_uniform(N, rs.random_sample());
    GR_CUDA_TRY(cudaMemcpyAsync(&centres[0], q->x + first, sizeof(double), cudaMemcpyDeviceToHost, st));
    GR_CUDA_TRY(cudaStreamSynchronize(st));
    pp_init_kernel<<<blocks, kRedThreads, 0, st>>>(q->x, N, centres[0], q->closest, q->partial);
    count_launch();
    double pot = 0.0;
    if (int rc = finish_to_host(q, blocks, 1, 1, &pot, st)) return rc;
    bool degenerate = false;
    const int nblk = (int)ceil_div<int64_t>(N, kPpBlock);
    int64_t* d_lo = q->small_i + 8;
    int64_t* d_hi = q->small_i + 16;
    double* d_pot = q->small + 48;
    std::vector<double> sorted_centres(1, centres[0]);
    sorted_centres.reserve((size_t)k);
    for (int c = 1; c < k; ++c) {
        for (int t = 0; t < trials; ++t) h[t] = rs.random_sample();
        GR_CUDA_TRY(cudaMemcpyAsync(d_rand, h, trials * sizeof(double), cudaMemcpyHostToDevice, st));
        GR_CUDA_TRY(cudaMemcpyAsync(q->centres_sorted, sorted_centres.data(),
                                    sorted_centres.size() * sizeof(double),
                                    cudaMemcpyHostToDevice, st));
        pp_block_sums_kernel<<<nblk, kRedThreads, 0, st>>>(q->closest, N, q->bsum);
        pp_scan_search_kernel<<<1, 1024, 0, st>>>(q->bsum, nblk, q->closest, q->x, q->xs, N, d_rand,
                                                  trials, q->centres_sorted,
                                                  (int)sorted_centres.size(), d_cand_i, d_cand_x,
                                                  d_lo, d_hi, d_pot);
        // the first centres' ranges are most of the data: stream; later ones: gather the range
        const bool stream_all = c < kPpStreamRounds;
        if (stream_all) {
            switch (trials) {
                case 2: pp_eval_all_kernel<2><<<blocks, kRedThreads, 0, st>>>(q->x, q->closest, N, d_cand_x, q->partial); break;
                case 3: pp_eval_all_kernel<3><<<blocks, kRedThreads, 0, st>>>(q->x, q->closest, N, d_cand_x, q->partial); break;
                case 4: pp_eval_all_kernel<4><<<blocks, kRedThreads, 0, st>>>(q->x, q->closest, N, d_cand_x, q->partial); break;
                case 5: pp_eval_all_kernel<5><<<blocks, kRedThreads, 0, st>>>(q->x, q->closest, N, d_cand_x, q->partial); break;
                case 6: pp_eval_all_kernel<6><<<blocks, kRedThreads, 0, st>>>(q->x, q->closest, N, d_cand_x, q->partial); break;
                case 7: pp_eval_all_kernel<7><<<blocks, kRedThreads, 0, st>>>(q->x, q->closest, N, d_cand_x, q->partial); break;
                default: pp_eval_all_kernel<8><<<blocks, kRedThreads, 0, st>>>(q->x, q->closest, N, d_cand_x, q->partial); break;
            }
        } else {
            dim3 eval_grid((unsigned)blocks, (unsigned)trials);
            pp_eval_range_kernel<<<eval_grid, kRedThreads, 0, st>>>(q->xs, q->perm, q->closest,
                                                                    d_cand_x, d_lo, d_hi, q->partial);
        }
        finish_sum_kernel<<<1, 32, 0, st>>>(q->partial, blocks, kMaxTrials, trials, q->small);
        count_launch(4);
        GR_CUDA_TRY(cudaGetLastError());
        // one copy for the doubles (gains [0..8), candidate values [32..40), pot [48]), one for
        // the ranges ([8..16) lo, [16..24) hi)
        double hd[49];
        int64_t hi64[24];
        GR_CUDA_TRY(cudaMemcpyAsync(hd, q->small, sizeof(hd), cudaMemcpyDeviceToHost, st));
        GR_CUDA_TRY(cudaMemcpyAsync(hi64, q->small_i, sizeof(hi64), cudaMemcpyDeviceToHost, st));
        GR_CUDA_TRY(cudaStreamSynchronize(st));
        const double* gains = hd;
        const double* cx = hd + 32;
        const int64_t* lo = hi64 + 8;
        const int64_t* hi = hi64 + 16;
        pot = hd[48];
        if (!(pot > 0.0)) { degenerate = true; break; }   // every distinct value is a centre already
        int best = 0;                                       // np.argmin: first minimum
        for (int t = 1; t < trials; ++t)
            if (pot - gains[t] < pot - gains[best]) best = t;
        centres[(size_t)c] = cx[best];
        sorted_centres.insert(std::upper_bound(sorted_centres.begin(), sorted_centres.end(), cx[best]),
                              cx[best]);
        if (stream_all) {
            pp_update_all_kernel<<<blocks, kRedThreads, 0, st>>>(q->x, N, cx[best], q->closest);
            count_launch();
        } else if (hi[best] > lo[best]) {
            const int ub = (int)std::min<int64_t>(blocks, ceil_div<int64_t>(hi[best] - lo[best],
                                                                            kRedThreads));
            pp_update_range_kernel<<<ub, kRedThreads, 0, st>>>(q->xs, q->perm, cx[best], lo[best],
                                                              hi[best], q->closest);
            count_launch();
        }
    }

    if (degenerate) {
        // fewer distinct values than bins: exact quantisation, every distinct value a level
        copy_matrix_kernel<T><<<blocks, 256, 0, st>>>(static_cast<const T*>(q->src), N, q->cols,
                                                      q->ld, ldo, out);
        count_runs_kernel<<<blocks, kRedThreads, 0, st>>>(q->xs, N, q->partial);
        count_launch(2);
        double runs = 0.0;
        if (int rc = finish_to_host(q, blocks, 1, 1, &runs, st)) return rc;
        if (n_distinct_out) *n_distinct_out = (int64_t)runs;
        if (n_iter_out) *n_iter_out = 0;
        if (centers_out)
            for (int c = 0; c < k; ++c) centers_out[c] = std::nan("");
        return GR_OK;
    }

    // ---- Lloyd loop, one launch (_kmeans.py:620-758) -------------------------------------------
    GR_CUDA_TRY(cudaMemcpyAsync(q->lloyd_centers, centres.data(), k * sizeof(double),
                                cudaMemcpyHostToDevice, st));
    GR_CUDA_TRY(cudaFuncSetAttribute(lloyd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(LloydShared)));
    const int threads = std::max(32, (k + 31) / 32 * 32);
    lloyd_kernel<<<1, threads, sizeof(LloydShared), st>>>(
        q->xs, q->P, N, k, q->lloyd_centers, q->var * tol, max_iter, q->mean, q->lloyd_info,
        q->lloyd_counts, q->thresholds, q->levels);
    count_launch();
    GR_CUDA_TRY(cudaGetLastError());
    int32_t info[4];
    GR_CUDA_TRY(cudaMemcpyAsync(info, q->lloyd_info, sizeof(info), cudaMemcpyDeviceToHost, st));
    GR_CUDA_TRY(cudaMemcpyAsync(centres.data(), q->lloyd_centers, k * sizeof(double),
                                cudaMemcpyDeviceToHost, st));
    GR_CUDA_TRY(cudaStreamSynchronize(st));
    const int n_used = info[1];
    quantize_kernel<T><<<blocks, 256, 0, st>>>(q->x, N, q->cols, ldo, q->thresholds, q->levels,
                                               n_used, out);
    count_launch();
    GR_CUDA_TRY(cudaGetLastError());
    if (n_distinct_out) {
        std::vector<double> lv((size_t)n_used);
        GR_CUDA_TRY(cudaMemcpyAsync(lv.data(), q->levels, n_used * sizeof(double),
                                    cudaMemcpyDeviceToHost, st));
        GR_CUDA_TRY(cudaStreamSynchronize(st));
        if (sizeof(T) == 4)
            for (auto& v : lv) v = (double)(float)v;      // what the output actually holds
        std::sort(lv.begin(), lv.end());
        *n_distinct_out = (int64_t)(std::unique(lv.begin(), lv.end()) - lv.begin());
    }
    if (n_iter_out) *n_iter_out = info[0];
    if (centers_out)
        for (int c = 0; c < k; ++c) centers_out[c] = centres[(size_t)c];
    return GR_OK;
}

template <typename T>
int kl_impl(const T* V, int64_t n, int32_t f, int64_t ldv, const T* G, int64_t ldg, const T* F,
            int64_t ldf, int32_t r, double* cost_out, int device, void* stream) {
    GR_REQUIRE(V && G && F && cost_out, "gr_mdl_error_cost: NULL argument");
    GR_REQUIRE(n >= 1 && f >= 1 && r >= 1 && r <= 64, "gr_mdl_error_cost: n/f/r = %lld/%d/%d (r <= 64)",
               (long long)n, f, r);
    GR_REQUIRE(ldv >= f && ldg >= r && ldf >= f, "gr_mdl_error_cost: row strides too small");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "cannot select device %d", device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    retain_async_pool(device);
    const int blocks = (int)std::min<int64_t>(n, kRedBlocks * 2);
    double* partial = nullptr;
    GR_CUDA_TRY(cudaMallocAsync(&partial, (blocks + 1) * sizeof(double), st));
    kl_cost_kernel<T><<<blocks, kRedThreads, 0, st>>>(V, n, f, ldv, G, ldg, F, ldf, r, partial);
    finish_sum_kernel<<<1, 32, 0, st>>>(partial, blocks, 1, 1, partial + blocks);
    count_launch(2);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(cost_out, partial + blocks, sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaFreeAsync(partial, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(GR_ERR_CUDA, "gr_mdl_error_cost: %s", cudaGetErrorString(e));
    return GR_OK;
}

template <typename T>
int kl_pair_impl(const T* V, const T* A, int64_t n, int32_t f, int64_t ldv, int64_t lda,
                 double* cost_out, int device, void* stream) {
    GR_REQUIRE(V && A && cost_out, "gr_mdl_kl: NULL argument");
    GR_REQUIRE(n >= 1 && f >= 1 && ldv >= f && lda >= f, "gr_mdl_kl: bad shape");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "cannot select device %d", device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = reduce_blocks(n * f);
    double* partial = nullptr;
    GR_CUDA_TRY(cudaMallocAsync(&partial, (blocks + 1) * sizeof(double), st));
    kl_pair_kernel<T><<<blocks, kRedThreads, 0, st>>>(V, A, n, f, ldv, lda, partial);
    finish_sum_kernel<<<1, 32, 0, st>>>(partial, blocks, 1, 1, partial + blocks);
    count_launch(2);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(cost_out, partial + blocks, sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaFreeAsync(partial, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(GR_ERR_CUDA, "gr_mdl_kl: %s", cudaGetErrorString(e));
    return GR_OK;
}

template <typename T>
int roles_impl(const T* W, int64_t n, int32_t r, int64_t ldw, int32_t* argmax, T* pct, int64_t ldp,
               int device, void* stream) {
    GR_REQUIRE(W != nullptr && n >= 1 && r >= 1 && ldw >= r, "gr_roles: bad factor matrix");
    GR_REQUIRE(argmax != nullptr || pct != nullptr, "gr_roles: both outputs are NULL");
    GR_REQUIRE(pct == nullptr || ldp >= r, "gr_roles: ldp < r");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "cannot select device %d", device);
    const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(n, 256), 148 * 8);
    roles_kernel<T><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(W, n, r, ldw, argmax, pct,
                                                                          ldp);
    count_launch();
    GR_CUDA_TRY(cudaGetLastError());
    return GR_OK;
}

}  // namespace

// ---- C ABI ------------------------------------------------------------------------------------
extern "C" int gr_quantizer_create(gr_quantizer_t** out, int64_t capacity, int device) {
    GR_REQUIRE(out != nullptr, "gr_quantizer_create: out is NULL");
    *out = nullptr;
    GR_REQUIRE(capacity >= 1, "gr_quantizer_create: capacity = %lld", (long long)capacity);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "gr_quantizer_create: cannot select device %d", device);
    if (int rc = require_sm100(device)) return rc;
    gr_quantizer* q = new (std::nothrow) gr_quantizer();
    if (!q) return fail(GR_ERR_OUT_OF_MEMORY, "gr_quantizer_create: host allocation failed");
    q->device = device;
    q->capacity = capacity;
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const double*)nullptr, (double*)nullptr,
                                    (const int64_t*)nullptr, (int64_t*)nullptr, capacity, 0, 64);
    cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (const double*)nullptr, (double*)nullptr,
                                  capacity);
    q->cub_temp_bytes = std::max(sort_bytes, scan_bytes) + 256;
    const size_t nb = (size_t)capacity * sizeof(double);
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    alloc((void**)&q->x, nb);
    alloc((void**)&q->xs, nb);
    alloc((void**)&q->P, nb + sizeof(double));
    alloc((void**)&q->closest, nb);
    alloc((void**)&q->perm, (size_t)capacity * sizeof(int64_t));
    alloc((void**)&q->bsum, ((size_t)ceil_div<int64_t>(capacity, kPpBlock) + 1) * sizeof(double));
    alloc((void**)&q->centres_sorted, kMaxBins * sizeof(double));
    alloc(&q->cub_temp, q->cub_temp_bytes);
    alloc((void**)&q->partial, (size_t)kRedBlocks * kMaxTrials * sizeof(double));
    alloc((void**)&q->small, 64 * sizeof(double));
    alloc((void**)&q->small_i, 32 * sizeof(int64_t));
    alloc((void**)&q->lloyd_centers, kMaxBins * sizeof(double));
    alloc((void**)&q->thresholds, kMaxBins * sizeof(double));
    alloc((void**)&q->levels, kMaxBins * sizeof(double));
    alloc((void**)&q->lloyd_info, 4 * sizeof(int32_t));
    alloc((void**)&q->lloyd_counts, kMaxBins * sizeof(int64_t));
    if (e != cudaSuccess) {
        cudaGetLastError();
        gr_quantizer_destroy(q);
        return fail(e == cudaErrorMemoryAllocation ? GR_ERR_OUT_OF_MEMORY : GR_ERR_CUDA,
                    "gr_quantizer_create: %s", cudaGetErrorString(e));
    }
    *out = q;
    return GR_OK;
}

extern "C" int gr_quantizer_destroy(gr_quantizer_t* q) {
    if (!q) return GR_OK;
    DeviceGuard guard(q->device);
    cudaFree(q->x);
    cudaFree(q->xs);
    cudaFree(q->P);
    cudaFree(q->closest);
    cudaFree(q->perm);
    cudaFree(q->bsum);
    cudaFree(q->centres_sorted);
    cudaFree(q->np_offs);
    cudaFree(q->np_lens);
    cudaFree(q->np_out);
    cudaFree(q->cub_temp);
    cudaFree(q->partial);
    cudaFree(q->small);
    cudaFree(q->small_i);
    cudaFree(q->lloyd_centers);
    cudaFree(q->thresholds);
    cudaFree(q->levels);
    cudaFree(q->lloyd_info);
    cudaFree(q->lloyd_counts);
    delete q;
    return GR_OK;
}

extern "C" int gr_quantizer_bind_f32(gr_quantizer_t* q, const float* X_dev, int64_t rows,
                                     int64_t cols, int64_t ld, void* stream) {
    return bind_impl<float>(q, X_dev, rows, cols, ld, stream);
}
extern "C" int gr_quantizer_bind_f64(gr_quantizer_t* q, const double* X_dev, int64_t rows,
                                     int64_t cols, int64_t ld, void* stream) {
    return bind_impl<double>(q, X_dev, rows, cols, ld, stream);
}

extern "C" int gr_quantizer_encode_f32(gr_quantizer_t* q, int32_t n_bins, uint32_t seed,
                                       int32_t max_iter, double tol, float* out_dev, int64_t ldo,
                                       double* centers_out_host, int32_t* n_iter_out,
                                       int64_t* n_distinct_out, void* stream) {
    return encode_impl<float>(q, n_bins, seed, max_iter, tol, out_dev, ldo, centers_out_host,
                              n_iter_out, n_distinct_out, stream);
}
extern "C" int gr_quantizer_encode_f64(gr_quantizer_t* q, int32_t n_bins, uint32_t seed,
                                       int32_t max_iter, double tol, double* out_dev, int64_t ldo,
                                       double* centers_out_host, int32_t* n_iter_out,
                                       int64_t* n_distinct_out, void* stream) {
    return encode_impl<double>(q, n_bins, seed, max_iter, tol, out_dev, ldo, centers_out_host,
                               n_iter_out, n_distinct_out, stream);
}

extern "C" int gr_quantizer_count_distinct(gr_quantizer_t* q, int64_t* n_distinct_out,
                                           void* stream) {
    GR_REQUIRE(q != nullptr && q->src != nullptr && n_distinct_out != nullptr,
               "gr_quantizer_count_distinct: no matrix is bound");
    DeviceGuard guard(q->device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "cannot select device %d", q->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = reduce_blocks(q->N);
    count_runs_kernel<<<blocks, kRedThreads, 0, st>>>(q->xs, q->N, q->partial);
    count_launch();
    double runs = 0.0;
    if (int rc = finish_to_host(q, blocks, 1, 1, &runs, st)) return rc;
    *n_distinct_out = (int64_t)runs;
    return GR_OK;
}

extern "C" int gr_mdl_error_cost_f32(const float* V, int64_t n, int32_t f, int64_t ldv,
                                     const float* G, int64_t ldg, const float* F, int64_t ldf,
                                     int32_t r, double* cost_out, int device, void* stream) {
    return kl_impl<float>(V, n, f, ldv, G, ldg, F, ldf, r, cost_out, device, stream);
}
extern "C" int gr_mdl_error_cost_f64(const double* V, int64_t n, int32_t f, int64_t ldv,
                                     const double* G, int64_t ldg, const double* F, int64_t ldf,
                                     int32_t r, double* cost_out, int device, void* stream) {
    return kl_impl<double>(V, n, f, ldv, G, ldg, F, ldf, r, cost_out, device, stream);
}
extern "C" int gr_mdl_kl_f64(const double* V, const double* V_approx, int64_t n, int32_t f,
                             int64_t ldv, int64_t lda, double* cost_out, int device, void* stream) {
    return kl_pair_impl<double>(V, V_approx, n, f, ldv, lda, cost_out, device, stream);
}

extern "C" int gr_roles_f32(const float* W, int64_t n, int32_t r, int64_t ldw, int32_t* argmax_dev,
                            float* pct_dev, int64_t ldp, int device, void* stream) {
    return roles_impl<float>(W, n, r, ldw, argmax_dev, pct_dev, ldp, device, stream);
}
extern "C" int gr_roles_f64(const double* W, int64_t n, int32_t r, int64_t ldw, int32_t* argmax_dev,
                            double* pct_dev, int64_t ldp, int device, void* stream) {
    return roles_impl<double>(W, n, r, ldw, argmax_dev, pct_dev, ldp, device, stream);
}

// NumPy RandomState(seed).random_sample(count): exposed so that the host tests can pin the stream
// the quantiser's seeding relies on (no device needed).
extern "C" int gr_numpy_random_sample(uint32_t seed, int64_t count, double* out_host) {
    GR_REQUIRE(out_host != nullptr && count >= 0, "gr_numpy_random_sample: bad arguments");
    NumpyRandomState rs(seed);
    for (int64_t i = 0; i < count; ++i) out_host[i] = rs.random_sample();
    return GR_OK;
}

// RandomState(seed).choice(n, p=np.full(n, 1 / n)): the first k-means++ centre (_kmeans.py:231).
extern "C" int gr_numpy_choice_uniform(uint32_t seed, int64_t n, int64_t* index_out) {
    GR_REQUIRE(index_out != nullptr && n >= 1, "gr_numpy_choice_uniform: bad arguments");
    NumpyRandomState rs(seed);
    *index_out = choice_uniform(n, rs.random_sample());
    return GR_OK;
}
