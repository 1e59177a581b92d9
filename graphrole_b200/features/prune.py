"""Feature pruning between recursion levels (host side).

Semantics follow graphrole/features/prune.py: every feature column is re-coded into vertical
logarithmic bins, columns whose binned versions differ by at most `feature_group_thresh`
everywhere (Chebyshev distance) are linked, and from each linked group only the oldest member
survives.  This is the caller-side bookkeeping of hot path A (SURVEY.md section 8, row A6 and
"next" #1), implemented with sort/inverse-index passes instead of the reference's per-bin
masking loop.
"""
from typing import Dict, Hashable, Iterable, List, Set

import numpy as np

from graphrole_b200.types import DataFrameDict, DataFrameLike, VectorLike


def vertical_log_binning(arr: VectorLike, frac: float = 0.5) -> np.ndarray:
    """Re-code `arr` into rank-based logarithmic bins.

    Bin 0 takes (about) the lowest `frac` of the values, bin 1 the same fraction of what is
    left, and so on; equal values never straddle a bin boundary and every bin holds at least
    one distinct value (cf. prune.py:13-56).
    """
    if not 0 < frac < 1:
        raise ValueError('must specify frac in interval (0, 1)')
    values = np.asarray(arr)
    total = values.shape[0]
    if total == 0:
        return np.zeros(0, dtype=int)
    _, inverse, counts = np.unique(values, return_inverse=True, return_counts=True)
    covered = np.cumsum(counts)           # covered[u] = #elements <= u-th distinct value
    bin_of_distinct = np.empty(covered.shape[0], dtype=int)
    done = 0                              # elements already assigned
    first = 0                             # first distinct value not yet assigned
    bin_id = 0
    while done < total:
        want = done + max(int(frac * (total - done)), 1)
        last = int(np.searchsorted(covered, want, side='left'))
        bin_of_distinct[first:last + 1] = bin_id
        done = int(covered[last])
        first = last + 1
        bin_id += 1
    return bin_of_distinct[inverse.reshape(-1)]


class _DisjointSets:
    """Union-find over column positions."""

    def __init__(self, size: int) -> None:
        self.parent = list(range(size))

    def find(self, i: int) -> int:
        root = i
        while self.parent[root] != root:
            root = self.parent[root]
        while self.parent[i] != root:
            self.parent[i], i = root, self.parent[i]
        return root

    def union(self, i: int, j: int) -> None:
        ri, rj = self.find(i), self.find(j)
        if ri != rj:
            self.parent[max(ri, rj)] = min(ri, rj)


class FeaturePruner:
    """Finds the redundant columns of a feature frame (cf. prune.py:59-139)."""

    def __init__(self, generation_dict: Dict[int, DataFrameDict], feature_group_thresh: int):
        self._generation_dict = generation_dict
        self._feature_group_thresh = feature_group_thresh

    def prune_features(self, features: DataFrameLike) -> List[Hashable]:
        """Names of the columns to drop: all but the oldest member of every group."""
        to_drop: List[Hashable] = []
        for group in self._group_features(features):
            if len(group) < 2:
                continue
            keep = self._get_oldest_feature(group)
            to_drop.extend(name for name in group if name != keep)
        return to_drop

    def _group_features(self, features: DataFrameLike) -> Iterable[Set[Hashable]]:
        """Connected components of the graph linking columns whose binned versions are within
        the threshold of each other in max-norm.  Columns linked to nothing are not reported
        (the reference's feature graph only has the endpoints of its edges as nodes)."""
        names = list(features.columns)
        n_feat = len(names)
        if n_feat < 2:
            return []
        gaps = self._binned_gaps(features)
        sets = _DisjointSets(n_feat)
        linked = np.zeros(n_feat, dtype=bool)
        for i in range(n_feat - 1):
            for j in np.nonzero(gaps[i, i + 1:] <= self._feature_group_thresh)[0]:
                sets.union(i, i + 1 + int(j))
                linked[i] = linked[i + 1 + int(j)] = True
        groups: Dict[int, Set[Hashable]] = {}
        for i in range(n_feat):
            if linked[i]:
                groups.setdefault(sets.find(i), set()).add(names[i])
        return list(groups.values())

    def _binned_gaps(self, features: DataFrameLike) -> np.ndarray:
        """[F, F] matrix of max-norm distances between the binned columns (host arithmetic;
        DeviceFeaturePruner computes the same integers on the GPU)."""
        names = list(features.columns)
        binned = np.stack([vertical_log_binning(features[name].to_numpy()) for name in names])
        gaps = np.zeros((len(names), len(names)), dtype=np.int64)
        for i in range(len(names) - 1):
            gap = np.abs(binned[i + 1:] - binned[i]).max(axis=1)
            gaps[i, i + 1:] = gap
            gaps[i + 1:, i] = gap
        return gaps

    def _get_oldest_feature(self, feature_names: Set[Hashable]) -> Hashable:
        """Member generated in the earliest generation; ties broken by sorted name."""
        for gen in range(len(self._generation_dict)):
            candidates = feature_names.intersection(self._generation_dict[gen].keys())
            if candidates:
                return self._set_getitem(candidates)
        return self._set_getitem(feature_names)

    @staticmethod
    def _set_getitem(s: Set[Hashable]) -> Hashable:
        """Deterministic pick from a set: its smallest element."""
        return min(s)


class DeviceFeaturePruner(FeaturePruner):
    """FeaturePruner whose O(n) work -- binning every column and the pairwise max-norm
    distances (prune.py:104-108) -- runs on the GPU (csrc/prune.cu through
    gr_prune_bin_f64 / gr_prune_pairwise_gap_i32); grouping and the oldest-member choice stay
    on the host.  Raises when the CUDA library or a device is missing."""

    def __init__(self, generation_dict: Dict[int, DataFrameDict], feature_group_thresh: int,
                 device=None) -> None:
        super().__init__(generation_dict, feature_group_thresh)
        self._device = device

    def _binned_gaps(self, features: DataFrameLike) -> np.ndarray:
        import torch
        from graphrole_b200 import _native
        device = torch.device(self._device if self._device is not None else 'cuda')
        values = np.ascontiguousarray(features.to_numpy(dtype=np.float64))
        pruner = _native.Pruner(values.shape[0], device)
        try:
            bins = pruner.bin_columns(torch.from_numpy(values).to(pruner.device))
            gaps = pruner.pairwise_gaps(bins).cpu().numpy().astype(np.int64)
        finally:
            pruner.close()
        return gaps
