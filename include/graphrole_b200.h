/*
 * graphrole_b200.h — C-ABI of libgraphrole_b200.so (sm_100a).
 *
 * The reference (dkaslovsky/GraphRole) is pure Python and has no FFI of its own; its two
 * data-parallel hot paths are Python-level call sites.  This header is the boundary a
 * maintainer of the reference would bind with ctypes to replace
 *
 *   path A  RecursiveFeatureExtractor._get_next_features
 *           graphrole/features/extract.py:98-119  (+ naming/layout step :144-163)
 *   path B  get_nmf_decomposition
 *           graphrole/roles/factor.py:10-26  ->  sklearn/decomposition/_nmf.py:726-888
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary; every entry point returns an int status
 *     (GR_OK == 0).  gr_last_error() returns a thread-local, NUL-terminated description of the
 *     most recent failure on the calling thread.
 *   - pointers named *_dev are device pointers owned by the caller (e.g. torch
 *     Tensor.data_ptr()); pointers named *_host are host pointers owned by the caller.  The
 *     library owns only the opaque handles it returns and the workspaces inside them.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Device-buffer
 *     entry points are asynchronous with respect to the host; *_host entry points synchronise
 *     the stream before returning.
 *   - one host thread per handle at a time; distinct handles may be used concurrently.
 */
#ifndef GRAPHROLE_B200_H
#define GRAPHROLE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    GR_OK = 0,
    GR_ERR_INVALID_ARGUMENT = 1,   /* bad shape / alignment / null pointer                   */
    GR_ERR_INVALID_GRAPH = 2,      /* rowptr not monotone, colidx out of range, ...           */
    GR_ERR_CUDA = 3,               /* a CUDA runtime call or kernel launch failed              */
    GR_ERR_UNSUPPORTED_DEVICE = 4, /* not an sm_100 device: there is no fallback path          */
    GR_ERR_OUT_OF_MEMORY = 5
};

/* ---- library ------------------------------------------------------------------------- */

const char* gr_last_error(void);
/* "graphrole_b200 <version> sm_100a" */
const char* gr_version(void);
/* Number of kernels this library has launched on the calling process since load
 * (monotone counter; bench.py differences it around the timed region). */
int64_t gr_kernel_launch_count(void);

/* ---- graph handle: the CSR form of graph.get_nodes()/get_neighbors() -------------------
 * Replaces the per-node neighbour lookups of
 *   graphrole/graph/interface/base.py:58-69, interface/networkx.py:36-46.
 * Row i of the CSR is the i-th node label in sorted order (base.py:24-25, extract.py:129);
 * colidx holds each row's UNIQUE out-neighbours (edge weights are ignored on this path,
 * a self loop contributes the node once).  Duplicate column indices inside a row are
 * accepted and counted as separate arcs.
 *
 * n_rows  rows held by this handle (a node-range shard holds a slice of the rows)
 * n_cols  number of rows of the feature matrices that colidx may address (global node count)
 * rowptr  int64[n_rows + 1], non-decreasing, rowptr[n_rows] == nnz; rowptr[0] is 0 for a whole
 *         graph and may be in (0, 32) for a row-range shard that keeps the colidx entries back to
 *         the previous 32-arc boundary: the kernel's summation order depends on arc offsets
 *         modulo 32, so such a shard reproduces the unsharded results bit for bit
 * colidx  int32[nnz], values in [0, n_cols)
 * The arrays are NOT copied: they must stay valid and unchanged until gr_csr_destroy().
 * flags:
 *   GR_CSR_VALIDATE   range-check colidx on the device (one extra pass over colidx)
 *   GR_CSR_HOT_HINTS  build a library-owned copy of colidx (4*nnz bytes) whose sign bit marks
 *                     arcs into the most frequently gathered rows (top rows by in-degree, as
 *                     many as fit a fixed share of L2); the gather kernel loads those rows
 *                     with an L2 evict_last policy.  Purely a cache hint: results are
 *                     bit-identical with and without it.
 * One stream per handle at a time: the handle's workspaces are reused by consecutive calls.
 */
typedef struct gr_csr gr_csr_t;

enum { GR_CSR_VALIDATE = 1, GR_CSR_HOT_HINTS = 2 };

int gr_csr_create(gr_csr_t** out, int64_t n_rows, int64_t n_cols, int64_t nnz,
                  const int64_t* rowptr_dev, const int32_t* colidx_dev,
                  int device, int flags);
int gr_csr_destroy(gr_csr_t* g);
/* Re-computes the GR_CSR_HOT_HINTS tags for feature rows of `row_bytes` bytes (default 256 =
 * 64 fp32 columns): the hot set is as many of the most gathered rows as fit the L2 budget, so a
 * column-group shard that aggregates 8 columns (32-byte rows) can pin 8 x as many rows.  Cache
 * hint only -- results do not change.  Synchronises the device. */
int gr_csr_tune_hot_rows(gr_csr_t* g, int64_t row_bytes);
/* n_rows, n_cols, nnz, number of hub rows (rows split across warps), number of hub segments,
 * number of rows tagged hot (0 without GR_CSR_HOT_HINTS) */
int gr_csr_info(const gr_csr_t* g, int64_t* n_rows, int64_t* n_cols, int64_t* nnz,
                int64_t* n_hub_rows, int64_t* n_hub_segments, int64_t* n_hot_rows);

/* ---- path A: one ReFeX recursion level -------------------------------------------------
 * Replaces graphrole/features/extract.py:105-118 (reindex -> agg([sum, mean]) -> fillna(0))
 * for all rows r in [row_lo, row_hi) of the handle:
 *     out_sum [r * ldo + c] = sum_{k in rowptr[r]..rowptr[r+1]} X[colidx[k] * ldx + c]
 *     out_mean[r * ldo + c] = out_sum / deg(r)            (0 when deg(r) == 0, extract.py:113)
 * for c in [0, d).  Column order of the pair (out_sum, out_mean) is the reference's
 * agg-major order (extract.py:158-162): pass out_mean = out_sum + d with ldo = 2*d to get the
 * reference's [c1(sum)..cd(sum), c1(mean)..cd(mean)] row layout in one buffer.
 * Either output pointer may be NULL to skip that aggregation.
 * fp32 storage; fp32 sums in 32/LPR independent lane-group accumulators per column (rows of at
 * most 2048 arcs), hub rows combined from 1024-arc partials in fp64.
 * X must not alias the outputs.
 */
int gr_refex_aggregate_f32(gr_csr_t* g, const float* X_dev, int64_t ldx, int32_t d,
                           int64_t row_lo, int64_t row_hi,
                           float* out_sum_dev, float* out_mean_dev, int64_t ldo,
                           void* stream);

/* ---- path A, node-range sharded across the GPUs of one box ----------
 * The reference has no distributed path (single-process Python; graphrole/features/extract.py
 * :77-87 simply calls _get_next_features once per level).  Sharded by output row, level l+1
 * needs ALL rows of level l's recursed block on every GPU -- the one exchange step of the
 * path.  gr_refex_aggregate_bcast_f32 fuses that exchange into the gather kernel: it is
 * gr_refex_aggregate_f32 whose mean rows are stored into n_replicas copies of the next input
 * matrix (this GPU's and its peers', mapped with gr_peer_open) instead of one out_mean buffer.
 *   mean_replicas  host array of n_replicas (<= 16) device pointers, each rebased to
 *                  handle-local row 0 (replica_base + shard_row_offset * ldo), row stride ldo
 *   out_sum        optional local sum rows (row stride ldo as well)
 * Results are bit-identical to gr_refex_aggregate_f32.  The caller orders levels across GPUs
 * with gr_peer_barrier (or any collective) on the same stream. */
int gr_refex_aggregate_bcast_f32(gr_csr_t* g, const float* X_dev, int64_t ldx, int32_t d,
                                 int64_t row_lo, int64_t row_hi, float* out_sum_dev,
                                 float* const* mean_replicas, int32_t n_replicas, int64_t ldo,
                                 void* stream);

/* Device buffers another process on the same box can map (cudaMalloc + CUDA IPC).
 * gr_peer_alloc   allocates `bytes` on `device` and fills the 64-byte handle to ship to the
 *                 other ranks (any byte transport: torch.distributed.all_gather_object, a file)
 * gr_peer_open    maps a peer's buffer into this process (peer access over NVLink is enabled on
 *                 first use); gr_peer_close unmaps it; gr_peer_free releases an own buffer. */
typedef struct { unsigned char bytes[64]; } gr_ipc_handle_t;

int gr_peer_alloc(void** dev_ptr_out, int64_t bytes, int device, gr_ipc_handle_t* handle_out);
int gr_peer_open(void** dev_ptr_out, const gr_ipc_handle_t* handle, int device);
int gr_peer_close(void* dev_ptr, int device);
int gr_peer_free(void* dev_ptr, int device);

/* Stream-ordered barrier across the ranks of a box through flags in peer-mapped memory.
 * flag_arrays[q] = rank q's flag array (gr_peer_flag_words() zero-initialised uint64 words in a
 * gr_peer_alloc buffer; the own one at index `rank`).  Every rank calls it with the same,
 * strictly increasing `epoch` (>= 1).  Work enqueued on `stream` after the call starts only
 * when every rank's work enqueued before its call has completed, peer stores included.
 * A rank that does not arrive within timeout_s (<= 0: 20 s) makes the kernel give up instead of
 * hanging the GPU; gr_peer_barrier_status then returns the epoch that timed out (0 = none). */
int gr_peer_barrier(void* const* flag_arrays, int32_t n_ranks, int32_t rank, int64_t epoch,
                    double timeout_s, void* stream);
int64_t gr_peer_flag_words(void);
int gr_peer_barrier_status(const void* own_flag_array, int64_t* timed_out_epoch);

/* Host-buffer variant (the call a ctypes binding inside the reference would make with
 * DataFrame.values): copies X (n_cols x d, row stride ldx) to the device, runs `levels`
 * recursion levels and copies every level's [n_rows, 2*d] (sum block | mean block) result
 * back.  Level l+1 aggregates the `recurse_on` block of level l (0 = sum, 1 = mean).
 * out_host: levels * n_rows * 2*d floats, level-major.  Synchronous. */
int gr_refex_levels_host_f32(gr_csr_t* g, const float* X_host, int64_t ldx, int32_t d,
                             int32_t levels, int32_t recurse_on, float* out_host, void* stream);

/* Host-buffer variant of a NODE-RANGE SHARD (section 8e): every rank of an exchange group calls
 * it with its shard handle.  X_host is the level-0 input of ALL n_cols nodes (row stride ldx; a
 * column group passes X_host + first_column); every rank copies only its own rows of it over
 * PCIe and pushes them to the peers' replicas over NVLink (the ranks' row ranges must tile
 * [0, n_cols)).  Each level is gr_refex_aggregate_bcast_f32 into
 * the other replica set followed by gr_peer_barrier; the OWN rows of every level are copied
 * back, overlapped with the next level: out_host [levels][2][n_rows][d] -- per level the sum
 * rows then the mean rows, each one contiguous block (the agg-major order of
 * extract.py:158-162 at block level; pitched copies of narrow rows waste the PCIe link).
 *   replicas_even / replicas_odd  n_ranks device pointers each: base of every rank's [n_cols, d]
 *                  replica used as input of the even / odd levels (own at index `rank`)
 *   flag_arrays, epoch_inout      as gr_peer_barrier; *epoch_inout is advanced levels + 1 times
 *   row_offset     global row number of the handle's row 0
 * Recurses on the mean block.  Synchronous.  n_ranks == 1 needs no peers (flags unused). */
int gr_refex_levels_host_sharded_f32(gr_csr_t* g, const float* X_host, int64_t ldx, int32_t d,
                                     int32_t levels, int64_t row_offset,
                                     float* const* replicas_even, float* const* replicas_odd,
                                     void* const* flag_arrays, int32_t n_ranks, int32_t rank,
                                     int64_t* epoch_inout, float* out_host, void* stream);

/* ---- path B: RolX NMF, multiplicative updates, Frobenius loss ---------------------------
 * Replaces sklearn.decomposition._nmf._fit_multiplicative_update (beta_loss=2, no
 * regularisation; _nmf.py:726-888) as called from graphrole/roles/factor.py:19-25.
 *   X  [n, f] row stride ldx, non-negative           W [n, r] row stride r (in: W0, out: W)
 *   H  [r, f] row stride f (in: H0, out: H)
 * Each iteration:  W *= (X H^T) / (W (H H^T));  H *= (W^T X) / ((W^T W) H)  with zero
 * denominators replaced by float32 eps (_nmf.py:32,615,701).  Every `check_every`
 * iterations (sklearn: 10) the Frobenius error ||X - W H||_F is evaluated and the loop stops
 * when (previous_error - error) / error_at_init < tol (_nmf.py:867-879).  tol == 0 disables
 * the test.  n_iter_out receives the number of iterations done, err_out the last evaluated
 * error (error at init when never evaluated).
 * Contractions run on tcgen05 tensor cores in TF32 with fp32 accumulation (use_tf32 != 0)
 * or on fp32 FFMA (use_tf32 == 0).  Asynchronous inputs, but the call synchronises the stream
 * at every convergence check (the stopping rule is data dependent).
 */
typedef struct gr_nmf gr_nmf_t;

int gr_nmf_create(gr_nmf_t** out, int64_t n, int32_t f, int32_t r, int device);
int gr_nmf_destroy(gr_nmf_t* h);
int gr_nmf_mu_f32(gr_nmf_t* h, const float* X_dev, int64_t ldx,
                  float* W_dev, float* H_dev,
                  int32_t max_iter, double tol, int32_t check_every, int32_t use_tf32,
                  int32_t* n_iter_out, double* err_out, void* stream);
/* 1 when the last gr_nmf_mu_f32 call on this handle ran the tcgen05 kernel, 0 for the FFMA
 * kernels (use_tf32 == 0, or a shape the tensor-core kernel does not take: it needs
 * f % 4 == 0, f <= 1024, ldx % 4 == 0 and a 16-byte aligned X; a rank that is not a multiple of
 * 4 runs there on zero-padded factors -- a zero role is a fixed point of the updates and adds
 * exact zeros to every sum -- so that every n_roles of the reference's default grid, 2..8,
 * takes the tensor-core path.  Feature counts that are not a multiple of 4: pad X with zero
 * columns, as graphrole_b200.roles.extract.DeviceModelGrid does). */
int gr_nmf_last_path(const gr_nmf_t* h);
/* 1 when gr_nmf_mu_f32(use_tf32 != 0) on this X would run the tcgen05 kernels, else 0. */
int gr_nmf_takes_tensor_cores(const gr_nmf_t* h, const float* X_dev, int64_t ldx);
/* Frobenius error ||X - W H||_F (dense-residual form, _nmf.py:122), fp64 accumulation. */
int gr_nmf_error_f32(gr_nmf_t* h, const float* X_dev, int64_t ldx,
                     const float* W_dev, const float* H_dev, double* err_out, void* stream);
/* The same quantity with W H formed on the tcgen05 tensor cores (W and H rounded to TF32, the
 * residual X - W H and its square in fp32, fp64 sums): the form gr_nmf_mu_f32 uses for its
 * convergence checks when use_tf32 != 0.  GR_ERR_INVALID_ARGUMENT for shapes the tensor-core
 * kernels do not take (see gr_nmf_last_path). */
int gr_nmf_error_tf32(gr_nmf_t* h, const float* X_dev, int64_t ldx,
                      const float* W_dev, const float* H_dev, double* err_out, void* stream);

/* Row-sharded form of path B: rows of X and W are split over the ranks, H
 * is replicated.  One multiplicative-update iteration is cut where the sums cross the ranks:
 *   gr_nmf_iteration_local_f32   W update of the local rows (in place) and the local sums
 *                                wtx_out [r, f] = W^T X, wtw_out [r, r] = W^T W (device buffers,
 *                                updated W); H is read only
 *   (caller: all-reduce(sum) of wtx / wtw over the ranks -- NCCL; r f + r^2 floats)
 *   gr_nmf_update_h_f32          H *= wtx / (wtw H) in place (_nmf.py:633-635, 701-721); identical
 *                                on every rank because its inputs are
 * Convergence: gr_nmf_error_tf32 / _f32 give the local ||X - W H||_F; the caller sums the
 * squares over the ranks.  Host loop: graphrole_b200/roles/sharded.py::RowShardedNmf. */
int gr_nmf_iteration_local_f32(gr_nmf_t* h, const float* X_dev, int64_t ldx, float* W_dev,
                               const float* H_dev, int32_t use_tf32, float* wtx_out_dev,
                               float* wtw_out_dev, void* stream);
int gr_nmf_update_h_f32(gr_nmf_t* h, const float* wtx_dev, const float* wtw_dev, float* H_dev,
                        void* stream);

/* ---- "next" rows of the path: what runs between the levels -------------
 *
 * Feature pruning, graphrole/features/prune.py.  The reference re-bins EVERY retained column
 * after every level (features.apply(vertical_log_binning), prune.py:104) and compares all column
 * pairs (pdist chebyshev, prune.py:107); connected components / oldest-member choice
 * (prune.py:94-130) stay on the host (F <= ~10^2 columns).
 *
 * gr_prune_bin_f32 / _f64   vertical_log_binning (prune.py:13-56) of every column of X
 *     [n_rows, d] (row stride ldx): bin 0 takes the lowest max(int(frac * n), 1) values plus
 *     everything tied with the largest of them, bin 1 the same share of what is left, ...
 *     Output is COLUMN-major: bins_dev[c * ldb + i] = bin of X[i, c] (ldb >= n_rows).
 *     Integer results, bit-exact with the reference for inputs without NaN (the extractor
 *     fills NaN with 0 first, extract.py:128-133); -0.0 and 0.0 tie like in np.unique.
 *     frac outside (0, 1) is refused like prune.py:19-20.
 * gr_prune_pairwise_gap_i32  gap_dev[i * d + j] = max_row |bins[i] - bins[j]| (Chebyshev
 *     distance of binned columns; the reference links i and j when it is <= the level's
 *     threshold, prune.py:108-111).  d <= 1024.  Exact.
 * The handle owns the sort workspaces for n_rows rows (about 160 bytes per row).
 */
typedef struct gr_pruner gr_pruner_t;

int gr_pruner_create(gr_pruner_t** out, int64_t n_rows, int device);
int gr_pruner_destroy(gr_pruner_t* h);
int gr_prune_bin_f32(gr_pruner_t* h, const float* X_dev, int64_t ldx, int32_t d, double frac,
                     int32_t* bins_dev, int64_t ldb, void* stream);
int gr_prune_bin_f64(gr_pruner_t* h, const double* X_dev, int64_t ldx, int32_t d, double frac,
                     int32_t* bins_dev, int64_t ldb, void* stream);
int gr_prune_pairwise_gap_i32(gr_pruner_t* h, const int32_t* bins_dev, int64_t ldb, int32_t d,
                              int32_t* gap_dev, void* stream);

/* Level-0 neighbourhood features from CSR arrays, replacing the per-node nx.ego_graph /
 * nx.edge_boundary loop of graphrole/graph/interface/networkx.py:48-83 (igraph.py:61-103).
 *   rowptr int64[n + 1], colidx int32[nnz] ascending and unique inside a row (out-neighbours),
 *   weights fp64[nnz] or NULL (= 1 per arc); an undirected graph stores every edge as two arcs
 *   and a self loop as one.
 * Outputs, fp64[n] each (exact for integer weights):
 *   out_weight  weighted out-degree                 in_weight  weighted in-degree (directed
 *   diag        weight of the node's self loop                 only; NULL otherwise)
 *   internal    `internal_edges`: weight inside the node's radius-1 (out-)egonet
 *   external    `external_edges`: weight leaving that egonet
 * Host composition (networkx.py:53-63): undirected `degree` = out_weight + diag (a self loop
 * counts twice, like nx.Graph.degree); directed in/out/total_degree = in, out, in + out. */
int gr_level0_features_f64(int64_t n, int64_t nnz, const int64_t* rowptr_dev,
                           const int32_t* colidx_dev, const double* weights_dev,
                           int32_t directed, double* out_weight_dev, double* in_weight_dev,
                           double* diag_dev, double* internal_dev, double* external_dev,
                           int device, void* stream);

/* ---- RolX epilogue: what RoleExtractor runs after the NMF ------------
 *
 * Lloyd-Max quantiser, replacing graphrole/roles/factor.py:29-49:
 *     KMeans(n_clusters=n_bins, random_state=seed).fit(X.reshape(X.size, 1)); centres[labels]
 * (sklearn/cluster/_kmeans.py: fit :1440-1563, k-means++ :180-278, Lloyd :620-758).
 * gr_quantizer_bind_*   takes the matrix (rows x cols, row stride ld, row-major flattening order
 *     like X.reshape): centres the entries by their mean -- summed in NumPy's pairwise order, so
 *     that it is X.mean() to the last bit (on grid-valued data the assignment of exactly
 *     equidistant points depends on it) --, sorts them once (values and permutation) and
 *     prefix-sums the sorted values -- shared by every n_bins tried on the same matrix (the grid
 *     of roles/extract.py:121-133 quantises one factor with 2^1 .. 2^8 bins).
 * gr_quantizer_encode_* k-means++ seeding with NumPy's RandomState(seed) stream (the reference
 *     hard-codes seed 1), Lloyd iterations until labels repeat or the squared centre shift is
 *     <= tol * var(X) (sklearn: tol 1e-4, max_iter 300), then out = centre of every entry's
 *     cluster (same shape, row stride ldo).  centers_out_host[n_bins] receives cluster_centers_ in
 *     sklearn's cluster order (NaN when the data hold fewer distinct values than n_bins -- then
 *     every distinct value is its own level and out == X), n_distinct_out the number of distinct
 *     output values (np.unique(encoded).size, description_length.py:37-38).
 *     n_bins > rows * cols is refused with sklearn's message "n_samples=... should be >=
 *     n_clusters=..." (GR_ERR_INVALID_ARGUMENT): callers rely on it (roles/extract.py:127-129).
 *     n_bins <= 1024.  Labels equal scikit-learn's when the data hold >= n_bins distinct values.
 * The handle owns about 50 bytes of device workspace per entry of `capacity`.
 */
typedef struct gr_quantizer gr_quantizer_t;

int gr_quantizer_create(gr_quantizer_t** out, int64_t capacity, int device);
int gr_quantizer_destroy(gr_quantizer_t* q);
int gr_quantizer_bind_f32(gr_quantizer_t* q, const float* X_dev, int64_t rows, int64_t cols,
                          int64_t ld, void* stream);
int gr_quantizer_bind_f64(gr_quantizer_t* q, const double* X_dev, int64_t rows, int64_t cols,
                          int64_t ld, void* stream);
int gr_quantizer_encode_f32(gr_quantizer_t* q, int32_t n_bins, uint32_t seed, int32_t max_iter,
                            double tol, float* out_dev, int64_t ldo, double* centers_out_host,
                            int32_t* n_iter_out, int64_t* n_distinct_out, void* stream);
int gr_quantizer_encode_f64(gr_quantizer_t* q, int32_t n_bins, uint32_t seed, int32_t max_iter,
                            double tol, double* out_dev, int64_t ldo, double* centers_out_host,
                            int32_t* n_iter_out, int64_t* n_distinct_out, void* stream);
/* np.unique(X).size of the bound matrix (description_length.py:37-38 on an encoded factor). */
int gr_quantizer_count_distinct(gr_quantizer_t* q, int64_t* n_distinct_out, void* stream);

/* Description-length error cost, replacing graphrole/roles/description_length.py:44-61 together
 * with the product at :19:  sum over V[i,j] != 0 of  v log(v / a) - v + a,  a = (G F)[i, j],
 * V [n, f], G [n, r], F [r, f] (r <= 64), fp64 arithmetic for either storage type, fixed
 * reduction order.  gr_mdl_kl_f64 is the same sum for an explicit approximation matrix
 * (get_error_cost(V, V_approx)).  Synchronous: cost_out is a host pointer. */
int gr_mdl_error_cost_f32(const float* V_dev, int64_t n, int32_t f, int64_t ldv,
                          const float* G_dev, int64_t ldg, const float* F_dev, int64_t ldf,
                          int32_t r, double* cost_out, int device, void* stream);
int gr_mdl_error_cost_f64(const double* V_dev, int64_t n, int32_t f, int64_t ldv,
                          const double* G_dev, int64_t ldg, const double* F_dev, int64_t ldf,
                          int32_t r, double* cost_out, int device, void* stream);
int gr_mdl_kl_f64(const double* V_dev, const double* V_approx_dev, int64_t n, int32_t f,
                  int64_t ldv, int64_t lda, double* cost_out, int device, void* stream);

/* RoleExtractor.roles / .role_percentage (graphrole/roles/extract.py:38-57): index of the first
 * maximum of every row of the node-role factor and the row divided by its sum.  Either output may
 * be NULL. */
int gr_roles_f32(const float* W_dev, int64_t n, int32_t r, int64_t ldw, int32_t* argmax_dev,
                 float* pct_dev, int64_t ldp, int device, void* stream);
int gr_roles_f64(const double* W_dev, int64_t n, int32_t r, int64_t ldw, int32_t* argmax_dev,
                 double* pct_dev, int64_t ldp, int device, void* stream);

/* np.random.RandomState(seed).random_sample(count) -- the stream the quantiser's seeding draws
 * from; host only, exposed so it can be pinned against NumPy without a device. */
int gr_numpy_random_sample(uint32_t seed, int64_t count, double* out_host);
/* RandomState(seed).choice(n, p=uniform): index of the first k-means++ centre (_kmeans.py:231);
 * exact up to n = 2^24, floor(u n) beyond (see csrc/rolx_epilogue.cu). */
int gr_numpy_choice_uniform(uint32_t seed, int64_t n, int64_t* index_out);

#ifdef __cplusplus
}
#endif
#endif /* GRAPHROLE_B200_H */
