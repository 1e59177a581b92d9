"""Level-0 ("generation 0") neighbourhood features from CSR arrays.

Host-side (scipy) closed forms of what the reference computes with a per-node
nx.ego_graph / nx.edge_boundary loop (graphrole/graph/interface/networkx.py:48-83).  This is
the input of the recursion, not one of the two accelerated hot paths (SURVEY.md section 8f,
"next" #2): it runs once per graph.

With W the weighted out-adjacency, B its 0/1 pattern and E = B or I the egonet membership
matrix (row i marks node i and its out-neighbours):

    T_i        = sum_{u in ego(i)} sum_{v in ego(i)} W[u, v]      = rowsum((E W) * E)
    internal_i = T_i                                  (directed: every arc inside the egonet)
               = (T_i + sum_{u in ego(i)} W[u, u]) / 2 (undirected: each edge once, loops once)
    external_i = sum_{u in ego(i)} outweight(u) - T_i  (arcs / edges leaving the egonet)
"""
import numpy as np
import pandas as pd
import scipy.sparse as sp

from graphrole_b200.graph.csr import CSRGraph


def _weighted_adjacency(g: CSRGraph) -> sp.csr_matrix:
    rp, ci = g.host_arrays()
    w = np.ones(g.nnz, dtype=np.float64) if g.weights is None else g.weights.cpu().numpy()
    return sp.csr_matrix((w, ci, rp), shape=(g.n, g.n))


def _maybe_int(values: np.ndarray, integral: bool):
    """The reference's columns are int64 when every edge weight is an int (networkx sums)."""
    if integral and np.all(values == np.round(values)):
        return values.astype(np.int64)
    return values


def local_degree_features(g: CSRGraph) -> pd.DataFrame:
    """Weighted degree columns: `degree` (undirected; a self loop counts twice, like
    nx.Graph.degree) or `in_degree`, `out_degree`, `total_degree` (directed)."""
    W = _weighted_adjacency(g)
    out_w = np.asarray(W.sum(axis=1)).ravel()
    index = list(g.node_labels())
    if g.directed:
        in_w = np.asarray(W.sum(axis=0)).ravel()
        cols = {'in_degree': in_w, 'out_degree': out_w, 'total_degree': in_w + out_w}
    else:
        cols = {'degree': out_w + W.diagonal()}
    return pd.DataFrame({k: _maybe_int(v, g.weights_integral) for k, v in cols.items()},
                        index=index)


def egonet_features(g: CSRGraph) -> pd.DataFrame:
    """`internal_edges` / `external_edges` of every node's radius-1 (out-)egonet."""
    W = _weighted_adjacency(g)
    n = g.n
    B = W.copy()
    B.data = np.ones_like(B.data)
    E = (B + sp.identity(n, format='csr', dtype=np.float64)).tocsr()
    E.data = np.ones_like(E.data)
    T = np.asarray((E @ W).multiply(E).sum(axis=1)).ravel()
    out_w = np.asarray(W.sum(axis=1)).ravel()
    ego_out = E @ out_w
    if g.directed:
        internal = T
    else:
        internal = (T + E @ W.diagonal()) / 2.0
    external = ego_out - T
    return pd.DataFrame({'internal_edges': _maybe_int(internal, g.weights_integral),
                         'external_edges': _maybe_int(external, g.weights_integral)},
                        index=list(g.node_labels()))


# ---- device path (csrc/level0.cu) -------------------------------------------------------------

def device_features(g: CSRGraph, device=None):
    """Level-0 columns computed on the GPU from the CSR arrays: ({column name: float64 CUDA
    tensor [n]}, in the reference's column order).  Same closed forms as above, one
    gr_level0_features_f64 call; the arrays are moved to `device` if they are not there yet."""
    import torch
    from graphrole_b200 import _native
    if device is None:
        device = g.rowptr.device if g.rowptr.is_cuda else torch.device('cuda')
    device = torch.device(device)
    weights = None if g.weights is None else g.weights.to(device)
    out = _native.level0_features(g.rowptr.to(device), g.colidx.to(device), weights,
                                  directed=g.directed)
    if g.directed:
        cols = {'in_degree': out['in_weight'], 'out_degree': out['out_weight'],
                'total_degree': out['in_weight'] + out['out_weight']}
    else:
        cols = {'degree': out['out_weight'] + out['diag']}
    cols['internal_edges'] = out['internal']
    cols['external_edges'] = out['external']
    return cols


def device_feature_frames(g: CSRGraph, device=None):
    """(local frame, egonet frame) like local_degree_features / egonet_features, computed by
    device_features."""
    cols = device_features(g, device)
    index = list(g.node_labels())
    host = {k: _maybe_int(v.cpu().numpy(), g.weights_integral) for k, v in cols.items()}
    ego = {k: host.pop(k) for k in ('internal_edges', 'external_edges')}
    return pd.DataFrame(host, index=index), pd.DataFrame(ego, index=index)
