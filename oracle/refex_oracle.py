"""ORACLE (test infrastructure, not product): CPU restatement of hot path A.

Restates RecursiveFeatureExtractor._get_next_features
(/root/reference/graphrole/features/extract.py:98-119, naming step :144-163) in float64.

Parity status: PINNED.  tests/test_oracle_refex.py checks every function here against the
fixtures in tests/golden/ that were produced by running the unmodified reference
(tests/golden/make_golden.py): the reference's own 4-node known-answer test
(tests/test_features/test_extract.py:104-122), its dangling-node case (:36-67), karate club
(examples/example.ipynb cell 3), directed / weighted / self-loop graphs and seeded random graphs
with injected float feature matrices.

Three restatements, from most literal to fastest:
  pandas_chain_rows   the reference's per-node pandas chain itself (used as the faithful CPU
                      timing port: this is what costs 1-50 ms per node in the reference)
  aggregate_loops     definitional pure-Python loops (tiny inputs only)
  aggregate_csr       S = A @ X, M = S / outdeg via SciPy CSR in float64 (scales to 10^7 nodes)
plus a C version (refex_oracle.c) used for sampled rows at full benchmark size.
"""
import ctypes
import os

import numpy as np
import pandas as pd
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
C_LIB_PATH = os.path.join(_HERE, 'librefex_oracle.so')


def aggregate_csr(rowptr, colidx, X):
    """(sum, mean) of X's rows over each CSR row's column indices, float64.

    extract.py:107-113: reindex(neighbours) -> agg([sum, mean]) -> fillna(0); a node without
    out-neighbours gets sum 0 and mean NaN -> 0.
    """
    rowptr = np.asarray(rowptr, dtype=np.int64)
    colidx = np.asarray(colidx, dtype=np.int64)
    X = np.asarray(X, dtype=np.float64)
    n_rows = rowptr.shape[0] - 1
    A = sp.csr_matrix((np.ones(colidx.shape[0]), colidx, rowptr), shape=(n_rows, X.shape[0]))
    S = np.asarray(A @ X)
    deg = np.diff(rowptr).astype(np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        M = np.where(deg[:, None] > 0, S / deg[:, None], 0.0)
    return S, M


def aggregate_loops(rowptr, colidx, X):
    """Definitional loops; for tiny inputs (cross-checks aggregate_csr)."""
    X = np.asarray(X, dtype=np.float64)
    n_rows = len(rowptr) - 1
    S = np.zeros((n_rows, X.shape[1]))
    M = np.zeros((n_rows, X.shape[1]))
    for i in range(n_rows):
        nbrs = [int(c) for c in colidx[rowptr[i]:rowptr[i + 1]]]
        for c in nbrs:
            S[i] += X[c]
        if nbrs:
            M[i] = S[i] / len(nbrs)
    return S, M


def next_features_frame(features, prev_features, nodes, neighbors, agg_names=('sum', 'mean')):
    """The DataFrame `_get_next_features` returns: rows in `nodes` order, columns agg-major
    named '<feature>(<agg>)' (extract.py:117-119,158-162).

    features: DataFrame indexed by node label; neighbors: callable node -> iterable of labels.
    """
    labels = list(features.index)
    row_of = {label: i for i, label in enumerate(labels)}
    nodes = list(nodes)
    rowptr = [0]
    colidx = []
    for node in nodes:
        colidx.extend(row_of[v] for v in neighbors(node))
        rowptr.append(len(colidx))
    X = features[list(prev_features)].to_numpy(dtype=np.float64)
    S, M = aggregate_csr(np.array(rowptr), np.array(colidx, dtype=np.int64), X)
    by_name = {'sum': S, 'mean': M}
    values = np.concatenate([by_name[a] for a in agg_names], axis=1)
    names = [f'{col}({a})' for a in agg_names for col in prev_features]
    return pd.DataFrame(values, index=nodes, columns=names)


def pandas_chain_rows(features, prev_features, rows, rowptr, colidx, aggs=('sum', 'mean')):
    """The reference's per-node chain (extract.py:105-118) executed for the row numbers in
    `rows` of a CSR graph whose row i is features.index[i].  Returns a DataFrame like
    `_get_next_features` restricted to those nodes.  This is the faithful CPU cost model of
    the reference (one reindex + agg + fillna + to_dict per node) used by bench.py."""
    labels = features.index
    prev_features = list(prev_features)
    aggs = list(aggs)
    out = {}
    for i in rows:
        nbr_labels = labels[colidx[rowptr[i]:rowptr[i + 1]]]
        agg = (features
               .reindex(index=nbr_labels, columns=prev_features)
               .agg(aggs)
               .fillna(0))
        flat = {}
        for agg_name, row in agg.to_dict(orient='index').items():
            for col, val in row.items():
                flat[f'{col}({agg_name})'] = val
        out[labels[i]] = flat
    return pd.DataFrame.from_dict(out, orient='index')


# ---- C restatement (refex_oracle.c) -------------------------------------------------------
_clib = None


def c_lib():
    """Load oracle/librefex_oracle.so (built by oracle/Makefile or __graft_entry__.build())."""
    global _clib
    if _clib is None:
        lib = ctypes.CDLL(C_LIB_PATH)
        lib.refex_oracle_rows_f32.restype = ctypes.c_int
        lib.refex_oracle_rows_f32.argtypes = [
            ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
        lib.refex_oracle_threads.restype = ctypes.c_int
        lib.refex_oracle_level_check_f64.restype = ctypes.c_int
        lib.refex_oracle_level_check_f64.argtypes = [
            ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
        _clib = lib
    return _clib


def aggregate_rows_c(rows, rowptr, colidx, X32, threads=0):
    """float64 (sum, mean) for the selected `rows` from a float32 X (the bytes the GPU sees),
    accumulated in float64 by the C restatement; `threads` = 0 uses every core."""
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    X32 = np.ascontiguousarray(X32, dtype=np.float32)
    d = X32.shape[1]
    S = np.empty((rows.shape[0], d), dtype=np.float64)
    M = np.empty((rows.shape[0], d), dtype=np.float64)
    rc = c_lib().refex_oracle_rows_f32(
        rows.shape[0], rows.ctypes.data, rowptr.ctypes.data, colidx.ctypes.data,
        X32.ctypes.data, X32.shape[1], d, S.ctypes.data, M.ctypes.data, threads)
    if rc != 0:
        raise RuntimeError(f'refex_oracle_rows_f32 failed with {rc}')
    return S, M


def c_threads():
    return int(c_lib().refex_oracle_threads())


def level_check_f64(rowptr, colidx, X64, gpu_out32, threads=0):
    """One level of the float64 recursion over the WHOLE graph, compared with the GPU's float32
    [sum | mean] output of the same level.  Returns (next-level float64 input = this level's
    means, max relative error of the sums, of the means, count of non-zero GPU entries where the
    float64 value is exactly 0).  X64 is what the reference would carry between levels
    (extract.py:77-83 never leaves float64)."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    X64 = np.ascontiguousarray(X64, dtype=np.float64)
    gpu_out32 = np.ascontiguousarray(gpu_out32, dtype=np.float32)
    n, d = X64.shape
    assert rowptr.shape[0] == n + 1 and gpu_out32.shape == (n, 2 * d)
    nxt = np.empty((n, d), dtype=np.float64)
    err = np.zeros(3, dtype=np.float64)
    rc = c_lib().refex_oracle_level_check_f64(
        n, rowptr.ctypes.data, colidx.ctypes.data, X64.ctypes.data, d, gpu_out32.ctypes.data,
        nxt.ctypes.data, err.ctypes.data, threads)
    if rc != 0:
        raise RuntimeError(f'refex_oracle_level_check_f64 failed with {rc}')
    return nxt, float(err[0]), float(err[1]), int(err[2])
