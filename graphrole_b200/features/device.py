"""Device-resident ReFeX: the whole `extract_features()` recursion with every O(n) step in HBM.

RecursiveFeatureExtractor (features/extract.py) keeps the reference's pandas-facing state
(`_features`, `_final_features`) and therefore moves every generation through host frames.  At
the benchmark sizes (10 M nodes) that round trip, not the kernels, is the cost, so this class
runs the same recursion on device tensors:

    level 0       gr_level0_features_f64          (networkx.py:48-83)
    aggregation   gr_refex_aggregate_f32          (extract.py:98-119)
    pruning       gr_prune_bin_f32 + gr_prune_pairwise_gap_i32, grouping on the host
                                                  (prune.py:13-56, 94-130; extract.py:121-142)

and only feature NAMES live on the host.  The rules are the reference's: generation g
aggregates the columns retained in generation g-1 (extract.py:104), candidates are appended to
all retained columns (extract.py:128-133), every column is binned and columns within Chebyshev
distance `g` of each other are grouped (extract.py:79-80, prune.py:104-111), the oldest member
of a group survives (prune.py:118-130), recursion stops when a generation retains nothing
(extract.py:86-87).  Values are fp32 (the aggregation kernel's type); binning sees exactly
those fp32 values, so the retained sets equal RecursiveFeatureExtractor's on the same device.
"""
from typing import Dict, Hashable, List, Optional

import numpy as np
import pandas as pd
import torch

from graphrole_b200 import _native
from graphrole_b200.features.prune import FeaturePruner
from graphrole_b200.graph import level0
from graphrole_b200.graph.csr import CSRGraph


class _Columns:
    """The `features.columns` view FeaturePruner needs."""

    def __init__(self, names: List[Hashable]):
        self.columns = list(names)


class _PrecomputedGapsPruner(FeaturePruner):
    def __init__(self, generation_dict, thresh, gaps: np.ndarray):
        super().__init__(generation_dict, thresh)
        self._gaps = gaps

    def _binned_gaps(self, features) -> np.ndarray:
        return self._gaps


class DeviceRecursiveFeatureExtractor:
    """ReFeX on a CSRGraph with features resident on one GPU.

    :param G: CSRGraph (host or device arrays; moved to `device`)
    :param max_generations: maximum levels of recursion (extract.py:26)
    :param device: CUDA device (default: the graph's device, else cuda:0)
    """

    aggs = ('sum', 'mean')

    def __init__(self, G: CSRGraph, max_generations: int = 10, device=None) -> None:
        if not isinstance(G, CSRGraph):
            raise TypeError('DeviceRecursiveFeatureExtractor takes a CSRGraph')
        if G.nnz == 0:
            raise ValueError('Input graph G must contain at least one edge')
        if device is None:
            device = G.rowptr.device if G.rowptr.is_cuda else torch.device('cuda', 0)
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _native.NativeLibraryError('graphrole_b200 has no CPU path')
        self.graph = G
        self.max_generations = max_generations
        self.generation_count = 0
        self._feature_group_thresh = 0
        self._names: List[str] = []                    # all retained columns, in frame order
        self._columns: Dict[str, torch.Tensor] = {}    # name -> fp32 [n]
        # binned columns of the working set, column-major and CONTIGUOUS in `_names` order (row k
        # of the buffer = column _names[k]): the pairwise kernel reads them in place, a
        # generation's candidates are binned straight into the rows behind the retained ones
        self._bins_buf: Optional[torch.Tensor] = None
        # values of every feature ever recorded as retained: a later generation may prune it from
        # the working set, the reference still reports it (its _final_features keeps the values)
        self._final_values: Dict[str, torch.Tensor] = {}
        self._final_features: Dict[int, List[str]] = {}
        self.timings_ms: Dict[str, float] = {}

    # ---- public ------------------------------------------------------------------------------
    def extract_features(self) -> pd.DataFrame:
        """Run the recursion; returns the retained features as a frame (latest generation's
        columns first, rows in sorted-label order) like the reference's extract_features()."""
        names, values = self.extract_features_device()
        return pd.DataFrame(values.cpu().numpy(), index=list(self.graph.node_labels()),
                            columns=names)

    def extract_features_device(self):
        """(column names, fp32 CUDA tensor [n, len(names)])."""
        if not self._final_features:
            self._run()
        names: List[str] = []
        for generation in reversed(list(self._final_features)):
            for name in self._final_features[generation]:
                if name not in names:
                    names.append(name)
        if not names:
            return names, torch.empty((self.graph.n, 0), device=self.device)
        return names, torch.stack([self._final_values[name] for name in names], dim=1)

    # ---- recursion -----------------------------------------------------------------------------
    def _run(self) -> None:
        g = self.graph
        self._handle = g.handle(self.device)
        self._pruner = _native.Pruner(g.n, self.device)
        try:
            with self._timed('level0'):
                cols = level0.device_features(g, self.device)
            self._update(list(cols), torch.stack([v.float() for v in cols.values()], dim=1))
            for generation in range(1, self.max_generations):
                self.generation_count = generation
                self._feature_group_thresh = generation
                names, values = self._get_next_features()
                self._update(names, values)
                if not self._final_features[generation]:
                    break
        finally:
            self._pruner.close()
            self._pruner = None

    def _get_next_features(self):
        """Candidates of this generation: [sum block | mean block] of the neighbours' columns
        retained in the previous generation, names `<feature>(<agg>)` agg-major."""
        prev = list(self._final_features[self.generation_count - 1])
        with self._timed('aggregate'):
            X = torch.stack([self._columns[name] for name in prev], dim=1).contiguous()
            out = self._handle.aggregate(X)
        names = [f'{col}({agg})' for agg in self.aggs for col in prev]
        return names, out

    def _update(self, names: List[str], values: torch.Tensor) -> None:
        """Append a generation's candidates, prune, record what was retained
        (extract.py:121-142)."""
        new = [(j, name) for j, name in enumerate(names) if name not in self._columns]
        k_old = len(self._names)
        need = k_old + len(names)
        if self._bins_buf is None or self._bins_buf.shape[0] < need:
            grown = torch.empty((max(need, 2 * k_old, 16), self.graph.n), dtype=torch.int32,
                                device=self.device)
            if self._bins_buf is not None and k_old:
                grown[:k_old].copy_(self._bins_buf[:k_old])
            self._bins_buf = grown
        with self._timed('bin'):
            if len(new) == len(names):
                self._pruner.bin_columns(values, out=self._bins_buf[k_old:k_old + len(names)])
            elif new:                                       # a re-derived name keeps its first values
                idx = torch.tensor([j for j, _ in new], device=self.device)
                self._pruner.bin_columns(values.index_select(1, idx).contiguous(),
                                         out=self._bins_buf[k_old:k_old + len(new)])
        for j, name in new:
            self._columns[name] = values[:, j]
            self._names.append(name)
        all_names = list(self._names)
        with self._timed('pairwise'):
            gaps = self._pruner.pairwise_gaps(self._bins_buf[:len(all_names)]).cpu().numpy()
        generations = {gen: dict.fromkeys(cols) for gen, cols in self._final_features.items()}
        pruner = _PrecomputedGapsPruner(generations, self._feature_group_thresh, gaps)
        redundant = set(pruner.prune_features(_Columns(all_names)))
        if redundant:
            with self._timed('compact'):
                keep = [k for k, name in enumerate(all_names) if name not in redundant]
                first = next(k for k, name in enumerate(all_names) if name in redundant)
                tail = [k for k in keep if k > first]
                # kept rows move one by one towards the front (row k lands at or before k, in
                # ascending order, so nothing is overwritten before it has been read)
                for dst, src in enumerate(tail, start=first):
                    self._bins_buf[dst].copy_(self._bins_buf[src])
            for name in redundant:
                self._names.remove(name)
                del self._columns[name]
        # columns.difference(redundant): sorted unique names (extract.py:140)
        # and pandas' Index.difference returns the names unsorted when nothing at all was dropped
        retained = list(pd.Index(names).difference(list(redundant)))
        self._final_features[self.generation_count] = retained
        for name in retained:
            self._final_values.setdefault(name, self._columns[name])

    # ---- timing --------------------------------------------------------------------------------
    class _Timer:
        def __init__(self, owner, key):
            self.owner, self.key = owner, key

        def __enter__(self):
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

        def __exit__(self, *exc):
            self.e1.record()
            self.e1.synchronize()
            t = self.owner.timings_ms
            t[self.key] = t.get(self.key, 0.0) + self.e0.elapsed_time(self.e1)

    def _timed(self, key):
        return self._Timer(self, key)
