// Level-0 ("generation 0") neighbourhood features straight from the CSR arrays
// ("next" #2).
//
// The reference builds them with one nx.ego_graph + nx.edge_boundary per node
// (graphrole/graph/interface/networkx.py:48-83; 3.9 ms per node measured).  With W the weighted
// out-adjacency and ego(i) = {i} U N_out(i):
//
//   out_w(i)   = sum_v W[i, v]            in_w(i) = sum_u W[u, i]          diag(i) = W[i, i]
//   T(i)       = sum_{u in ego(i)} sum_{v in ego(i)} W[u, v]
//              = out_w(i) + sum_{u in N(i), u != i} <row u restricted to ego(i)>
//   internal_i = T(i)                                      directed  (arcs inside the egonet)
//              = (T(i) + sum_{u in ego(i)} diag(u)) / 2    undirected (each edge once)
//   external_i = sum_{u in ego(i)} out_w(u) - T(i)
//
// (the closed forms of graphrole_b200/graph/level0.py, checked there against the reference's
// golden level-0 tables).  The restricted row sums are sorted-list intersections: for the arc
// (i, u) the shorter of row u / row i is walked by an 8-lane group and each entry is looked up in
// the longer one by binary search, so an arc costs min(deg) * log(max deg) -- hubs do not square.
// Eight lanes per row, four rows per warp; all sums in fp64 with a fixed reduction order (exact
// for integer weights); only in_w uses atomics.  Integer / index work, bound by L2 latency of the
// dependent binary-search loads, not by HBM.

#include <cub/device/device_scan.cuh>

#include <cstdlib>

#include "common.cuh"

using namespace gr;

namespace {

constexpr int kLanes = 8;        // lanes per arc group
constexpr int kBigRow = 1024;    // rows at least this long get a whole CTA

__device__ __forceinline__ double group_sum(double v, unsigned mask) {
#pragma unroll
    for (int o = kLanes / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, kLanes);
    return v;
}

// position of `v` in colidx[lo, hi) (ascending) or -1
__device__ __forceinline__ int64_t find(const int32_t* __restrict__ colidx, int64_t lo, int64_t hi,
                                        int32_t v) {
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const int32_t c = colidx[mid];
        if (c < v) lo = mid + 1; else hi = mid;
    }
    return lo;   // caller checks bounds + equality
}

__global__ void row_stats_kernel(int64_t n, const int64_t* __restrict__ rowptr,
                                 const int32_t* __restrict__ colidx,
                                 const double* __restrict__ w, double* __restrict__ out_w,
                                 double* __restrict__ in_w, double* __restrict__ diag) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = gid / kLanes;
    const int sub = (int)(gid % kLanes);
    const unsigned mask = 0xffffffffu;
    const bool valid = i < n;
    const int64_t a = valid ? rowptr[i] : 0, b = valid ? rowptr[i + 1] : 0;
    double s = 0.0, dg = 0.0;
    for (int64_t k = a + sub; k < b; k += kLanes) {
        const double wk = w ? w[k] : 1.0;
        const int32_t v = colidx[k];
        s += wk;
        if (v == i) dg += wk;
        if (in_w) atomicAdd(in_w + v, wk);
    }
    s = group_sum(s, mask);
    dg = group_sum(dg, mask);
    if (valid && sub == 0) {
        out_w[i] = s;
        diag[i] = dg;
    }
}

// Partial sums of one row's egonet terms: this lane is lane `sub` of arc-group `grp` out of
// `n_grp` groups working on row i (group g takes arcs g, g + n_grp, ...).
struct EgoPartial { double t, eo, ed; };

__device__ __forceinline__ EgoPartial ego_partial(int64_t i, int sub, int grp, int n_grp,
                                                  const int64_t* __restrict__ rowptr,
                                                  const int32_t* __restrict__ colidx,
                                                  const double* __restrict__ w,
                                                  const double* __restrict__ out_w,
                                                  const double* __restrict__ diag) {
    EgoPartial r = {0.0, 0.0, 0.0};
    const int64_t ia = rowptr[i], ib = rowptr[i + 1];
    const int32_t self = (int32_t)i;
    // is i its own neighbour (self loop)?
    const int64_t ps = find(colidx, ia, ib, self);
    const bool has_self = ps < ib && colidx[ps] == self;
    for (int64_t k = ia + grp; k < ib; k += n_grp) {
        const int32_t u = colidx[k];
        if (u == self) continue;                 // the u = i term is out_w(i), added by the caller
        const int64_t ua = rowptr[u], ub = rowptr[u + 1];
        if (sub == 0) {
            r.eo += out_w[u];
            r.ed += diag[u];
        }
        if (ub - ua <= ib - ia) {
            // walk row u, look every entry up in ego(i)
            for (int64_t q = ua + sub; q < ub; q += kLanes) {
                const int32_t v = colidx[q];
                bool in_ego = v == self;
                if (!in_ego) {
                    const int64_t p = find(colidx, ia, ib, v);
                    in_ego = p < ib && colidx[p] == v;
                }
                if (in_ego) r.t += w ? w[q] : 1.0;
            }
        } else {
            // walk ego(i) = row i (+ i itself when there is no self loop), look up in row u
            for (int64_t q = ia + sub; q < ib; q += kLanes) {
                const int32_t v = colidx[q];
                const int64_t p = find(colidx, ua, ub, v);
                if (p < ub && colidx[p] == v) r.t += w ? w[p] : 1.0;
            }
            if (!has_self && sub == 0) {
                const int64_t p = find(colidx, ua, ub, self);
                if (p < ub && colidx[p] == self) r.t += w ? w[p] : 1.0;
            }
        }
    }
    return r;
}

__device__ __forceinline__ void ego_finish(int64_t i, double t, double eo, double ed,
                                           const double* __restrict__ out_w,
                                           const double* __restrict__ diag, int directed,
                                           double* __restrict__ internal,
                                           double* __restrict__ external) {
    const double T = out_w[i] + t;
    internal[i] = directed ? T : (T + diag[i] + ed) * 0.5;
    external[i] = out_w[i] + eo - T;
}

// rows shorter than kBigRow: one 8-lane group per row; longer rows are appended to big_rows
__global__ void egonet_kernel(int64_t n, const int64_t* __restrict__ rowptr,
                              const int32_t* __restrict__ colidx, const double* __restrict__ w,
                              const double* __restrict__ out_w, const double* __restrict__ diag,
                              int directed, double* __restrict__ internal,
                              double* __restrict__ external, int32_t* __restrict__ big_rows,
                              int32_t* __restrict__ n_big) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = gid / kLanes;
    const int sub = (int)(gid % kLanes);
    EgoPartial r = {0.0, 0.0, 0.0};
    bool mine = i < n;
    if (mine && rowptr[i + 1] - rowptr[i] >= kBigRow) {
        if (sub == 0) big_rows[atomicAdd(n_big, 1)] = (int32_t)i;
        mine = false;
    }
    if (mine) r = ego_partial(i, sub, 0, 1, rowptr, colidx, w, out_w, diag);
    const double t = group_sum(r.t, 0xffffffffu);
    if (mine && sub == 0) ego_finish(i, t, r.eo, r.ed, out_w, diag, directed, internal, external);
}

// long rows: one 256-thread CTA (32 groups of 8 lanes) per row, fixed-order block reduction
__global__ void __launch_bounds__(256)
egonet_big_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                  const double* __restrict__ w, const double* __restrict__ out_w,
                  const double* __restrict__ diag, int directed, double* __restrict__ internal,
                  double* __restrict__ external, const int32_t* __restrict__ big_rows,
                  const int32_t* __restrict__ n_big) {
    __shared__ double part[3][256 / kLanes];
    const int grp = threadIdx.x / kLanes, sub = threadIdx.x % kLanes;
    const int count = *n_big;
    for (int idx = blockIdx.x; idx < count; idx += gridDim.x) {
        const int64_t i = big_rows[idx];
        const EgoPartial r = ego_partial(i, sub, grp, 256 / kLanes, rowptr, colidx, w, out_w, diag);
        const double t = group_sum(r.t, 0xffffffffu);
        __syncthreads();
        if (sub == 0) {
            part[0][grp] = t;
            part[1][grp] = r.eo;
            part[2][grp] = r.ed;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double st = 0.0, so = 0.0, sd = 0.0;
            for (int g = 0; g < 256 / kLanes; ++g) {
                st += part[0][g];
                so += part[1][g];
                sd += part[2][g];
            }
            ego_finish(i, st, so, sd, out_w, diag, directed, internal, external);
        }
    }
}


// ---- fast path: undirected, unweighted, no self loops (the benchmark graphs) -----------------------
// There the egonet terms collapse to triangle counts:
//     internal(i) = deg(i) + tri(i)            external(i) = sum_{u in N(i)} deg(u) - deg(i) - 2 tri(i)
// and triangles are enumerated ONCE each on the degree-oriented graph: arc u -> v is kept when
// (deg u, u) < (deg v, v), a triangle u < v < w is found as w in N+(u) ^ N+(v) and credited to its
// three corners with integer atomics (exact, order independent).  Oriented lists are short even
// where the graph has hubs -- a hub is late in the order, so it keeps almost none of its arcs -- and
// the walk costs sum over oriented arcs of min(|N+(u)|, |N+(v)|) log max(...), ~10 x less than
// intersecting the full rows (C3: 324 -> see profiles/ ms).
__device__ __forceinline__ bool precedes(int64_t du, int32_t u, int64_t dv, int32_t v) {
    return du < dv || (du == dv && u < v);
}

// degrees as 32-bit words: the orientation looks up the degree of every neighbour (nnz random
// reads); 4 n bytes stay in L2 where the two 8-byte row pointers of an 8 (n + 1)-byte array do not
__global__ void degree_kernel(int64_t n, const int64_t* __restrict__ rowptr,
                              int32_t* __restrict__ deg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) deg[i] = (int32_t)(rowptr[i + 1] - rowptr[i]);
}

__global__ void orient_count_kernel(int64_t n, const int64_t* __restrict__ rowptr,
                                    const int32_t* __restrict__ colidx,
                                    const int32_t* __restrict__ deg,
                                    int32_t* __restrict__ out_count,
                                    unsigned long long* __restrict__ nbr_deg_sum) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = gid / kLanes;
    const int sub = (int)(gid % kLanes);
    const bool valid = i < n;
    const int64_t a = valid ? rowptr[i] : 0, b = valid ? rowptr[i + 1] : 0;
    const int64_t di = b - a;
    int cnt = 0;
    unsigned long long dsum = 0;
    for (int64_t k = a + sub; k < b; k += kLanes) {
        const int32_t v = colidx[k];
        const int64_t dv = deg[v];
        dsum += (unsigned long long)dv;
        cnt += precedes(di, (int32_t)i, dv, v) ? 1 : 0;
    }
#pragma unroll
    for (int o = kLanes / 2; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o, kLanes);
        dsum += __shfl_xor_sync(0xffffffffu, dsum, o, kLanes);
    }
    if (valid && sub == 0) {
        out_count[i] = cnt;
        nbr_deg_sum[i] = dsum;
    }
}

// one thread per row: copy the kept arcs in order (rows are ascending, so the oriented rows are too)
__global__ void orient_fill_kernel(int64_t n, const int64_t* __restrict__ rowptr,
                                   const int32_t* __restrict__ colidx,
                                   const int32_t* __restrict__ deg,
                                   const int64_t* __restrict__ optr, int32_t* __restrict__ ocol) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = gid / kLanes;
    const int sub = (int)(gid % kLanes);
    if (i >= n) return;
    const int64_t a = rowptr[i], b = rowptr[i + 1], di = b - a;
    int64_t dst = optr[i];
    // the 8 lanes take 8 consecutive arcs per step; ranks inside the step by ballot
    const unsigned group_mask = 0xffu << ((threadIdx.x & 31) / kLanes * kLanes);
    for (int64_t k0 = a; k0 < b; k0 += kLanes) {
        const int64_t k = k0 + sub;
        bool keep = false;
        int32_t v = 0;
        if (k < b) {
            v = colidx[k];
            keep = precedes(di, (int32_t)i, deg[v], v);
        }
        // only the row's own 8 lanes vote: the other groups of the warp have other trip counts
        const unsigned m = __ballot_sync(group_mask, keep) & group_mask;
        const unsigned below = m & ((1u << (threadIdx.x & 31)) - 1u);
        if (keep) ocol[dst + __popc(below)] = v;
        dst += __popc(m);
    }
}

// 8-lane group per row u: for every v in N+(u) intersect N+(u) and N+(v).
// N+(u) of up to kHashMax arcs (every row of a preferential-attachment graph: a node keeps the
// arcs to its higher-degree neighbours) goes into a 64-slot open-addressing table in shared
// memory; the lanes then stream N+(v) -- coalesced -- and probe the table: one or two
// shared-memory reads per element instead of a binary search of ~log2|N+| dependent loads.
// Longer rows, and neighbours whose list is much longer than u's, keep the binary search over
// the shorter list.
constexpr int kHashSlots = 64, kHashMax = 40;
__device__ __forceinline__ int hash_slot(int32_t w) {
    return (int)(((uint32_t)w * 2654435761u) >> 26);          // 6 bits
}

__global__ void __launch_bounds__(256)
triangle_kernel(int64_t n, const int64_t* __restrict__ optr, const int32_t* __restrict__ ocol,
                unsigned long long* __restrict__ tri) {
    __shared__ int32_t table[256 / kLanes][kHashSlots];
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t u = gid / kLanes;
    const int sub = (int)(gid % kLanes);
    const int grp = threadIdx.x / kLanes;
    if (u >= n) return;                            // whole groups leave together
    const unsigned group_mask = 0xffu << ((threadIdx.x & 31) / kLanes * kLanes);
    const int64_t ua = optr[u], ub = optr[u + 1];
    const bool hashed = ub > ua && ub - ua <= kHashMax;
    if (hashed) {
        for (int s = sub; s < kHashSlots; s += kLanes) table[grp][s] = -1;
        __syncwarp(group_mask);
        for (int64_t q = ua + sub; q < ub; q += kLanes) {
            const int32_t w = ocol[q];
            int h = hash_slot(w);
            while (atomicCAS(&table[grp][h], -1, w) != -1) h = (h + 1) & (kHashSlots - 1);
        }
        __syncwarp(group_mask);
    }
    unsigned long long mine = 0;                   // triangles credited to u by this lane
    for (int64_t k = ua; k < ub; ++k) {
        const int32_t v = ocol[k];
        const int64_t va = optr[v], vb = optr[v + 1];
        unsigned long long found = 0;              // credited to v by this lane
        if (hashed && vb - va <= 4 * (ub - ua)) {
            for (int64_t q = va + sub; q < vb; q += kLanes) {
                const int32_t w = ocol[q];
                int h = hash_slot(w);
                int32_t t;
                while ((t = table[grp][h]) != -1) {
                    if (t == w) { ++found; atomicAdd(tri + w, 1ull); break; }
                    h = (h + 1) & (kHashSlots - 1);
                }
            }
        } else if (vb - va <= ub - ua) {
            // rows are sorted by node id, not by the orientation order: search the whole other row
            for (int64_t q = va + sub; q < vb; q += kLanes) {
                const int32_t w = ocol[q];
                const int64_t p = find(ocol, ua, ub, w);
                if (p < ub && ocol[p] == w) { ++found; atomicAdd(tri + w, 1ull); }
            }
        } else {
            for (int64_t q = ua + sub; q < ub; q += kLanes) {
                const int32_t w = ocol[q];
                const int64_t p = find(ocol, va, vb, w);
                if (p < vb && ocol[p] == w) { ++found; atomicAdd(tri + w, 1ull); }
            }
        }
        if (found) atomicAdd(tri + v, found);
        mine += found;
    }
    if (mine) atomicAdd(tri + u, mine);
}

__global__ void triangle_finish_kernel(int64_t n, const int64_t* __restrict__ rowptr,
                                       const unsigned long long* __restrict__ tri,
                                       const unsigned long long* __restrict__ nbr_deg_sum,
                                       double* __restrict__ out_w, double* __restrict__ diag,
                                       double* __restrict__ internal, double* __restrict__ external) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double deg = (double)(rowptr[i + 1] - rowptr[i]), t = (double)tri[i];
    out_w[i] = deg;
    diag[i] = 0.0;
    internal[i] = deg + t;
    external[i] = (double)nbr_deg_sum[i] - deg - 2.0 * t;
}

__global__ void any_self_loop_kernel(int64_t n, const int64_t* __restrict__ rowptr,
                                     const int32_t* __restrict__ colidx, int* __restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t a = rowptr[i], b = rowptr[i + 1];
    const int64_t p = find(colidx, a, b, (int32_t)i);
    if (p < b && colidx[p] == (int32_t)i) *flag = 1;
}

// returns GR_OK and sets *done = 1 when the fast path applied
int level0_triangles(int64_t n, int64_t nnz, const int64_t* rowptr, const int32_t* colidx,
                     double* out_w, double* diag, double* internal, double* external,
                     cudaStream_t st, int* done) {
    *done = 0;
    if (nnz == 0 || nnz >= ((int64_t)1 << 32)) return GR_OK;
    const unsigned blocks_n = (unsigned)ceil_div<int64_t>(n, 256);
    const unsigned blocks_g = (unsigned)ceil_div<int64_t>(n * kLanes, 256);
    // workspace: flag | counts (int32 n + 1) | optr (int64 n + 1) | nbr_deg_sum (u64 n) |
    // tri (u64 n) | ocol (int32 nnz: every arc is kept at most once) | deg (int32 n) | cub temp
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int32_t*)nullptr, (int64_t*)nullptr,
                                  n + 1);
    char* ws = nullptr;
    const size_t o_flag = 0, o_cnt = 256, o_optr = o_cnt + ((size_t)(n + 1) * 4 + 255) / 256 * 256,
                 o_dsum = o_optr + ((size_t)(n + 1) * 8 + 255) / 256 * 256,
                 o_tri = o_dsum + ((size_t)n * 8 + 255) / 256 * 256,
                 o_ocol = o_tri + ((size_t)n * 8 + 255) / 256 * 256,
                 o_deg = o_ocol + ((size_t)nnz * 4 + 255) / 256 * 256,
                 o_tmp = o_deg + ((size_t)n * 4 + 255) / 256 * 256,
                 total = o_tmp + scan_bytes + 256;
    GR_CUDA_TRY(cudaMallocAsync((void**)&ws, total, st));
    int* flag = reinterpret_cast<int*>(ws + o_flag);
    int32_t* cnt = reinterpret_cast<int32_t*>(ws + o_cnt);
    int64_t* optr = reinterpret_cast<int64_t*>(ws + o_optr);
    unsigned long long* dsum = reinterpret_cast<unsigned long long*>(ws + o_dsum);
    unsigned long long* tri = reinterpret_cast<unsigned long long*>(ws + o_tri);
    int32_t* ocol = reinterpret_cast<int32_t*>(ws + o_ocol);
    int32_t* deg = reinterpret_cast<int32_t*>(ws + o_deg);
    cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(int), st);
    any_self_loop_kernel<<<blocks_n, 256, 0, st>>>(n, rowptr, colidx, flag);
    int h_flag = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    count_launch();
    if (e != cudaSuccess || h_flag) {           // self loops: the general kernels handle them
        cudaFreeAsync(ws, st);
        if (e != cudaSuccess) return fail(GR_ERR_CUDA, "level-0: %s", cudaGetErrorString(e));
        return GR_OK;
    }
    cudaMemsetAsync(cnt + n, 0, sizeof(int32_t), st);
    cudaMemsetAsync(tri, 0, (size_t)n * sizeof(unsigned long long), st);
    degree_kernel<<<blocks_n, 256, 0, st>>>(n, rowptr, deg);
    orient_count_kernel<<<blocks_g, 256, 0, st>>>(n, rowptr, colidx, deg, cnt, dsum);
    size_t tmp = scan_bytes;
    e = cub::DeviceScan::ExclusiveSum(ws + o_tmp, tmp, cnt, optr, n + 1, st);
    orient_fill_kernel<<<blocks_g, 256, 0, st>>>(n, rowptr, colidx, deg, optr, ocol);
    triangle_kernel<<<blocks_g, 256, 0, st>>>(n, optr, ocol, tri);
    triangle_finish_kernel<<<blocks_n, 256, 0, st>>>(n, rowptr, tri, dsum, out_w, diag, internal,
                                                     external);
    count_launch(6);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFreeAsync(ws, st);
    if (e != cudaSuccess) return fail(GR_ERR_CUDA, "level-0 triangle path: %s", cudaGetErrorString(e));
    *done = 1;
    return GR_OK;
}

}  // namespace

extern "C" int gr_level0_features_f64(int64_t n, int64_t nnz, const int64_t* rowptr_dev,
                                      const int32_t* colidx_dev, const double* weights_dev,
                                      int32_t directed, double* out_weight_dev,
                                      double* in_weight_dev, double* diag_dev,
                                      double* internal_dev, double* external_dev, int device,
                                      void* stream) {
    GR_REQUIRE(n >= 0 && n < ((int64_t)1 << 31) && nnz >= 0,
               "gr_level0_features_f64: n = %lld, nnz = %lld", (long long)n, (long long)nnz);
    if (n == 0) return GR_OK;
    GR_REQUIRE(rowptr_dev != nullptr && (colidx_dev != nullptr || nnz == 0) && out_weight_dev != nullptr && diag_dev != nullptr &&
               internal_dev != nullptr && external_dev != nullptr,
               "gr_level0_features_f64: NULL argument");
    GR_REQUIRE(!directed || in_weight_dev != nullptr,
               "gr_level0_features_f64: a directed graph needs in_weight_dev");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "gr_level0_features_f64: cannot select device %d", device);
    if (int rc = require_sm100(device)) return rc;
    retain_async_pool(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // undirected + unweighted (+ no self loop, checked on the device): triangle counting on the
    // degree-oriented graph; GR_LEVEL0_GENERAL=1 forces the general kernels
    if (!directed && weights_dev == nullptr && !getenv("GR_LEVEL0_GENERAL")) {
        int done = 0;
        if (int rc = level0_triangles(n, nnz, rowptr_dev, colidx_dev, out_weight_dev, diag_dev,
                                      internal_dev, external_dev, st, &done))
            return rc;
        if (done) return GR_OK;
    }
    double* in_w = directed ? in_weight_dev : nullptr;
    if (in_w) GR_CUDA_TRY(cudaMemsetAsync(in_w, 0, (size_t)n * sizeof(double), st));
    const unsigned blocks = (unsigned)ceil_div<int64_t>(n * kLanes, 256);
    row_stats_kernel<<<blocks, 256, 0, st>>>(n, rowptr_dev, colidx_dev, weights_dev,
                                             out_weight_dev, in_w, diag_dev);
    // rows of >= kBigRow arcs are collected by the first kernel and finished by the second
    int32_t* big = nullptr;
    GR_CUDA_TRY(cudaMallocAsync((void**)&big, (size_t)(nnz / kBigRow + 2) * sizeof(int32_t), st));
    cudaMemsetAsync(big, 0, sizeof(int32_t), st);            // big[0] = counter, list from big + 1
    egonet_kernel<<<blocks, 256, 0, st>>>(n, rowptr_dev, colidx_dev, weights_dev, out_weight_dev,
                                          diag_dev, directed, internal_dev, external_dev, big + 1,
                                          big);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    egonet_big_kernel<<<4 * sms, 256, 0, st>>>(rowptr_dev, colidx_dev, weights_dev, out_weight_dev,
                                               diag_dev, directed, internal_dev, external_dev,
                                               big + 1, big);
    cudaFreeAsync(big, st);
    count_launch(3);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return fail(GR_ERR_CUDA, "level-0 kernels failed to launch: %s", cudaGetErrorString(e));
    return GR_OK;
}
