"""CSRGraph: the array form of the reference's graph plugin calls.

The reference walks `graph.get_nodes()` / `graph.get_neighbors(node)`
(graphrole/graph/interface/base.py:58-69) once per node per recursion level.  Here the same
information is flattened once into CSR arrays that live in HBM:

    rowptr  int64[n + 1]
    colidx  int32[nnz]      unique out-neighbours of row i, ascending
    row i   <->  i-th node label in sorted order (base.py:24-25, extract.py:129)

Edge weights never enter the recursion (extract.py:107-111 gathers rows, it does not scale
them); they are kept only because level-0 features are weighted (networkx.py:48-83).
"""
from typing import Callable, Hashable, Iterable, Optional, Sequence

import numpy as np
import torch


class CSRGraph:
    """Compressed sparse row out-adjacency with optional node labels and arc weights."""

    def __init__(self, rowptr, colidx, labels: Optional[Sequence[Hashable]] = None,
                 directed: bool = False, weights=None, weights_integral: bool = True):
        self.rowptr = _as_tensor(rowptr, torch.int64)
        self.colidx = _as_tensor(colidx, torch.int32)
        if self.rowptr.dim() != 1 or self.rowptr.numel() < 1:
            raise ValueError('rowptr must be a 1-D array of length n + 1')
        self.n = self.rowptr.numel() - 1
        self.n_cols = self.n  # rows of the feature matrix colidx may address (> n for a shard)
        self.nnz = self.colidx.numel()
        self.directed = bool(directed)
        self.labels = None if labels is None else list(labels)
        if self.labels is not None and len(self.labels) != self.n:
            raise ValueError('labels must have one entry per row')
        self.weights = None if weights is None else _as_tensor(weights, torch.float64)
        if self.weights is not None and self.weights.numel() != self.nnz:
            raise ValueError('weights must have one entry per arc')
        self.weights_integral = bool(weights_integral)
        self._handles = {}
        self._device_copies = {}

    # ---- construction ---------------------------------------------------------------
    @classmethod
    def from_edges(cls, src, dst, n: Optional[int] = None, directed: bool = False, weights=None,
                   labels: Optional[Sequence[Hashable]] = None,
                   weights_integral: bool = True) -> 'CSRGraph':
        """Build from integer edge arrays (row numbers, not labels).

        Undirected edges are stored as two arcs (a self loop as one).  Repeated edges collapse
        to one arc -- the reference's neighbour sets are sets (networkx.py:46) -- keeping the
        last weight, like re-adding an edge to a NetworkX graph does.
        """
        src = np.asarray(src, dtype=np.int64).ravel()
        dst = np.asarray(dst, dtype=np.int64).ravel()
        if src.shape != dst.shape:
            raise ValueError('src and dst must have the same length')
        if n is None:
            n = int(max(src.max(initial=-1), dst.max(initial=-1)) + 1)
        if src.size and (min(src.min(), dst.min()) < 0 or max(src.max(), dst.max()) >= n):
            raise ValueError('edge endpoint outside [0, n)')
        w = None if weights is None else np.asarray(weights, dtype=np.float64).ravel()
        if not directed:
            # an undirected edge is one object whichever way it is written: collapse repeats on
            # the unordered pair first (last weight wins for BOTH directions), then mirror
            lo, hi = np.minimum(src, dst), np.maximum(src, dst)
            order = np.argsort(lo * np.int64(n) + hi, kind='stable')
            pair = (lo * np.int64(n) + hi)[order]
            last = np.ones(pair.size, dtype=bool)
            last[:-1] = pair[1:] != pair[:-1]
            lo, hi = lo[order][last], hi[order][last]
            if w is not None:
                w = w[order][last]
            loop = lo == hi
            src, dst = np.concatenate([lo, hi[~loop]]), np.concatenate([hi, lo[~loop]])
            if w is not None:
                w = np.concatenate([w, w[~loop]])
        key = src * np.int64(n) + dst
        order = np.argsort(key, kind='stable')
        key = key[order]
        # keep the LAST occurrence of each (row, col)
        last = np.ones(key.size, dtype=bool)
        last[:-1] = key[1:] != key[:-1]
        key = key[last]
        if w is not None:
            w = w[order][last]
        rows = key // n
        rowptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.bincount(rows, minlength=n), out=rowptr[1:])
        return cls(rowptr, (key % n).astype(np.int32), labels=labels, directed=directed,
                   weights=w, weights_integral=weights_integral)

    @classmethod
    def from_neighbors(cls, nodes: Iterable[Hashable],
                       neighbors: Callable[[Hashable], Iterable[Hashable]],
                       directed: bool = True) -> 'CSRGraph':
        """Generic ingest through the reference's plugin calls: `nodes` is what
        graph.get_nodes() yields and `neighbors(node)` what graph.get_neighbors(node) yields.
        Rows are numbered in sorted-label order."""
        labels = sorted(nodes)
        row_of = {label: i for i, label in enumerate(labels)}
        rowptr = np.zeros(len(labels) + 1, dtype=np.int64)
        cols = []
        for i, label in enumerate(labels):
            nbrs = sorted({row_of[v] for v in neighbors(label)})
            cols.extend(nbrs)
            rowptr[i + 1] = len(cols)
        return cls(rowptr, np.asarray(cols, dtype=np.int32), labels=labels, directed=directed)

    @classmethod
    def from_edges_device(cls, src: torch.Tensor, dst: torch.Tensor, n: int,
                          directed: bool = False) -> 'CSRGraph':
        """Same as from_edges (unweighted) but with torch ops on the tensors' device, for graphs
        too large for host-side Python objects (10 M nodes / 200 M edges)."""
        src = src.to(torch.int64)
        dst = dst.to(torch.int64)
        if not directed:
            keep = src != dst
            key = torch.cat([src * n + dst, dst[keep] * n + src[keep]])
        else:
            key = src * n + dst
        del src, dst
        key = torch.unique(key, sorted=True)
        rows = torch.div(key, n, rounding_mode='floor')
        colidx = (key - rows * n).to(torch.int32)
        del key
        counts = torch.bincount(rows, minlength=n)
        del rows
        rowptr = torch.zeros(n + 1, dtype=torch.int64, device=colidx.device)
        torch.cumsum(counts, 0, out=rowptr[1:])
        return cls(rowptr, colidx, directed=directed)

    # ---- wire formats (the reference has none: graph objects only) ---
    def save(self, path: str) -> None:
        """Write the arrays (and labels / weights when present) as one .npz file."""
        rp, ci = self.host_arrays()
        arrays = {'rowptr': rp, 'colidx': ci, 'directed': np.array(self.directed),
                  'n_cols': np.array(self.n_cols),
                  'weights_integral': np.array(self.weights_integral)}
        if self.weights is not None:
            arrays['weights'] = self.weights.cpu().numpy()
        if self.labels is not None:
            arrays['labels'] = np.asarray(self.labels)
        np.savez(path, **arrays)

    @classmethod
    def load(cls, path: str, device=None) -> 'CSRGraph':
        """Read a file written by save(); `device` moves the arrays there."""
        with np.load(path, allow_pickle=False) as z:
            labels = z['labels'].tolist() if 'labels' in z.files else None
            g = cls(z['rowptr'], z['colidx'], labels=labels, directed=bool(z['directed']),
                    weights=z['weights'] if 'weights' in z.files else None,
                    weights_integral=bool(z['weights_integral']))
            g.n_cols = int(z['n_cols'])
        if device is not None:
            g = g.to(device)
        return g

    @classmethod
    def from_edge_list_file(cls, path: str, directed: bool = False, weighted: bool = False,
                            comment: str = '#', delimiter: Optional[str] = None) -> 'CSRGraph':
        """Text edge list, one `u v [w]` per line (whitespace or `delimiter` separated, lines
        starting with `comment` skipped).  Node tokens become labels; rows are numbered in
        sorted-label order (integers sort numerically), like every other ingest path."""
        import pandas as pd
        cols = [0, 1, 2] if weighted else [0, 1]
        sep = delimiter if delimiter is not None else r'\s+'
        frame = pd.read_csv(path, sep=sep, comment=comment, header=None, usecols=cols,
                            engine='python' if delimiter is None else 'c', dtype={0: str, 1: str})
        u, v = frame[0].to_numpy(), frame[1].to_numpy()
        try:
            u, v = u.astype(np.int64), v.astype(np.int64)
        except ValueError:
            pass
        labels, inverse = np.unique(np.concatenate([u, v]), return_inverse=True)
        src, dst = inverse[:u.size], inverse[u.size:]
        w = frame[2].to_numpy(dtype=np.float64) if weighted else None
        integral = bool(w is None or np.all(w == np.round(w)))
        return cls.from_edges(src, dst, n=labels.size, directed=directed, weights=w,
                              labels=labels.tolist(), weights_integral=integral)

    # ---- queries mirroring the plugin API ---------------------------------------------
    def node_labels(self):
        return self.labels if self.labels is not None else range(self.n)

    def num_edges(self) -> int:
        """Edge count with the meaning of NetworkX's number_of_edges()."""
        if self.directed:
            return self.nnz
        rp, ci = self.host_arrays()
        rows = np.repeat(np.arange(self.n), np.diff(rp))
        loops = int(np.count_nonzero(rows == ci))
        return (self.nnz - loops) // 2 + loops

    def out_degree(self) -> torch.Tensor:
        return self.rowptr[1:] - self.rowptr[:-1]

    def host_arrays(self):
        return self.rowptr.cpu().numpy(), self.colidx.cpu().numpy()

    # ---- device residency -------------------------------------------------------------
    def to(self, device) -> 'CSRGraph':
        device = torch.device(device)
        if self.rowptr.device == device:
            return self
        key = str(device)
        if key not in self._device_copies:
            g = CSRGraph(self.rowptr.to(device), self.colidx.to(device), labels=self.labels,
                         directed=self.directed,
                         weights=None if self.weights is None else self.weights.to(device),
                         weights_integral=self.weights_integral)
            g.n_cols = self.n_cols
            self._device_copies[key] = g
        return self._device_copies[key]

    def handle(self, device=None):
        """Native gr_csr_t handle for this graph on `device` (created once, cached)."""
        from graphrole_b200 import _native
        if device is None:
            device = self.rowptr.device if self.rowptr.is_cuda else torch.device('cuda', 0)
        device = torch.device(device)
        if device.type != 'cuda':
            raise _native.NativeLibraryError(
                'graphrole_b200 runs its hot paths on CUDA devices only (no CPU fallback)')
        if device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        key = str(device)
        if key not in self._handles:
            g = self.to(device)
            self._handles[key] = _native.CsrHandle(g.rowptr, g.colidx, n_cols=self.n_cols)
        return self._handles[key]

    def row_slice(self, lo: int, hi: int) -> 'CSRGraph':
        """Rows [lo, hi) as a shard; colidx still addresses all n nodes.  The colidx slice starts
        at the 32-arc boundary at or below the first arc (rowptr[0] of the shard is that
        remainder, < 32): the gather kernel reads colidx in 32-entry chunks aligned to absolute
        arc positions and its fp32 summation order follows the chunking, so a shard laid out
        this way gives bit-identical results to the unsharded graph."""
        rp = self.rowptr[lo:hi + 1]
        a, b = int(rp[0]), int(rp[-1])
        a0 = a - a % 32
        shard = CSRGraph(rp - a0, self.colidx[a0:b], directed=True)
        shard.n_cols = self.n
        shard.nnz = b - a            # arcs of the shard's rows (the slice holds a - a0 more)
        return shard

    def __repr__(self):
        return (f'CSRGraph(n={self.n}, nnz={self.nnz}, directed={self.directed}, '
                f'device={self.rowptr.device})')


def _as_tensor(a, dtype):
    if isinstance(a, torch.Tensor):
        return a.to(dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a), dtype=_NP[dtype]))


_NP = {torch.int64: np.int64, torch.int32: np.int32, torch.float64: np.float64}
