"""TEST INFRASTRUCTURE -- CPU oracle for the RolX epilogue.

Restates, in plain NumPy, what the reference does after the NMF:

  encode                 graphrole/roles/factor.py:29-49 -- Lloyd-Max quantiser = 1-D
                         KMeans(n_clusters=n_bins, random_state=1) on the flattened matrix.  The
                         arithmetic lives in scikit-learn (requirements.txt:4 `>=1.3.1`, installed
                         1.9.0): k-means++ seeding sklearn/cluster/_kmeans.py:180-278, Lloyd loop
                         :620-758 (+ _k_means_lloyd.pyx / _k_means_common.pyx for the E/M step,
                         empty-cluster relocation and the centre shift), fit driver :1440-1563.
  get_encoding_cost      graphrole/roles/description_length.py:32-41
  get_error_cost         graphrole/roles/description_length.py:44-61
  grid / rescale         graphrole/roles/extract.py:98-142, 163-173

Pinned by tests/test_oracle_rolx.py against tests/golden/rolx_cases.npz, which
tests/golden/make_golden.py produced by running the unmodified reference (and with it the
installed scikit-learn).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this module; the product never does.
"""
import numpy as np


def _sq_dist_to(c, x, x_sq):
    """sklearn.metrics.pairwise._euclidean_distances(c[None], X, Y_norm_squared=x_sq,
    squared=True) for one feature: -2 x.c + c.c + x.x, clipped at 0 (pairwise.py:377-412)."""
    d = -2.0 * (c * x)
    d += c * c
    d += x_sq
    np.maximum(d, 0, out=d)
    return d


def kmeans_plusplus_1d(x, k, rs):
    """_kmeans_plusplus (_kmeans.py:180-278) on centred 1-D data with unit sample weights.
    Returns the indices of the chosen points."""
    n = x.size
    x_sq = x * x
    trials = 2 + int(np.log(k))
    w = np.ones(n)
    idx = np.full(k, -1, dtype=np.int64)
    idx[0] = rs.choice(n, p=w / w.sum())
    closest = _sq_dist_to(x[idx[0]], x, x_sq)
    pot = closest @ w
    for c in range(1, k):
        rand_vals = rs.uniform(size=trials) * pot
        cand = np.searchsorted(np.cumsum(w * closest), rand_vals)
        np.clip(cand, None, n - 1, out=cand)
        dist = np.stack([np.minimum(closest, _sq_dist_to(x[j], x, x_sq)) for j in cand])
        pots = dist @ w
        best = int(np.argmin(pots))
        pot, closest, idx[c] = pots[best], dist[best], cand[best]
    return idx


def lloyd_1d(x, centers, tol, max_iter=300):
    """_kmeans_single_lloyd (_kmeans.py:620-758) for one feature, one thread.
    Returns (labels, centers, n_iter)."""
    n, k = x.size, centers.size
    centers = centers.copy()
    labels_old = np.full(n, -1, dtype=np.int64)
    strict = False
    it = 0
    for it in range(max_iter):
        # E step: argmin_j (|c_j|^2 - 2 x c_j), first minimum wins (_k_means_lloyd.pyx)
        score = centers[None, :] ** 2 - 2.0 * x[:, None] * centers[None, :]
        labels = np.argmin(score, axis=1)
        sums = np.bincount(labels, weights=x, minlength=k)
        counts = np.bincount(labels, minlength=k).astype(float)
        empty = np.where(counts == 0)[0]
        if empty.size:
            # _relocate_empty_clusters_dense (_k_means_common.pyx): the points farthest from
            # their centre become the empty clusters' centres
            dist = (x - centers[labels]) ** 2
            far = np.argpartition(dist, -empty.size)[:-empty.size - 1:-1]
            for new_id, far_idx in zip(empty, far):
                old_id = labels[far_idx]
                sums[old_id] -= x[far_idx]
                sums[new_id] = x[far_idx]
                counts[new_id] = 1
                counts[old_id] -= 1
        new_centers = centers.copy()
        nz = counts > 0
        new_centers[nz] = sums[nz] * (1.0 / counts[nz])
        shift = np.abs(new_centers - centers)
        centers = new_centers
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if (shift ** 2).sum() <= tol:
            break
        labels_old = labels
    if not strict:
        score = centers[None, :] ** 2 - 2.0 * x[:, None] * centers[None, :]
        labels = np.argmin(score, axis=1)
    return labels, centers, it + 1


def numpy_pairwise_sum(a):
    """np.add.reduce of a contiguous float64 vector, restated (numpy/_core/src/umath/
    loops_utils.h.src, pairwise sum): blocks of <= 128 values with eight running accumulators
    combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail; longer inputs are
    split at n/2 rounded down to a multiple of 8 and the halves added.  The quantiser's bind sums
    the data mean in exactly this order (csrc/rolx_epilogue.cu: numpy_leaf_sums_kernel +
    numpy_combine) because on grid-valued data scikit-learn's assignment of exactly-equidistant
    points falls with the last bit of X.mean(); tests/test_oracle_rolx.py pins this restatement
    to the installed NumPy bit for bit."""
    a = np.asarray(a, dtype=np.float64)
    n = a.size
    if n < 8:
        r = 0.0
        for x in a:
            r += x
        return r
    if n <= 128:
        r = a[:8].copy()
        m = n - (n % 8)
        for i in range(8, m, 8):
            r += a[i:i + 8]
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        for i in range(m, n):
            res += a[i]
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return numpy_pairwise_sum(a[:n2]) + numpy_pairwise_sum(a[n2:])


def kmeans_1d(values, k, seed=1, tol=1e-4, max_iter=300):
    """KMeans(n_clusters=k, random_state=seed).fit(values.reshape(-1, 1)) (_kmeans.py:1440-1563):
    returns (labels, cluster_centers_, n_iter_)."""
    v = np.asarray(values, dtype=np.float64).ravel()
    if v.size < k:
        raise ValueError(f'n_samples={v.size} should be >= n_clusters={k}.')
    rs = np.random.RandomState(seed)
    mean = v.mean()
    x = v - mean
    abs_tol = np.var(v) * tol
    seeds = kmeans_plusplus_1d(x, k, rs)
    labels, centers, n_iter = lloyd_1d(x, x[seeds], abs_tol, max_iter)
    return labels, centers + mean, n_iter


def encode(X, n_bins, seed=1):
    """graphrole/roles/factor.py:29-49: every entry replaced by its cluster centre."""
    X = np.asarray(X, dtype=np.float64)
    labels, centers, _ = kmeans_1d(X.reshape(X.size), n_bins, seed)
    return centers[labels].reshape(X.shape)


def encoding_cost(G_encoded, F_encoded):
    """description_length.py:32-41."""
    n_bins = max(len(np.unique(G_encoded)), len(np.unique(F_encoded)))
    return np.ceil(np.log2(n_bins)) * (G_encoded.size + F_encoded.size)


def error_cost(V, V_approx):
    """description_length.py:44-61: sum over v != 0 of v log(v / v') - v + v'."""
    total = 0.0
    for v, a in zip(np.asarray(V, dtype=float).ravel(), np.asarray(V_approx, dtype=float).ravel()):
        if v != 0:
            with np.errstate(divide='ignore'):
                total += v * np.log(v / a) - v + a
    return total


def rescale_costs(costs):
    """roles/extract.py:163-173."""
    norms = np.sqrt(np.nansum(np.square(costs), axis=1))
    with np.errstate(invalid='ignore', divide='ignore'):
        return costs / norms.reshape(costs.shape[0], 1)


def select_model(enc_costs, err_costs):
    """roles/extract.py:135-141: (roles, bits) of the smallest rescaled total cost."""
    total = rescale_costs(enc_costs) + rescale_costs(err_costs)
    r, b = np.argwhere(total == np.nanmin(total))[0]
    return int(r), int(b)
