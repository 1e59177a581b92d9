"""ReFeX recursive feature extraction with the neighbourhood aggregation on the GPU.

API surface of graphrole/features/extract.py (class, constructor, attributes and method
names, error behaviour) with `_get_next_features` -- the reference's per-node pandas loop,
extract.py:98-119 -- replaced by one CSR gather-reduce kernel launch per recursion level
(graphrole_b200/csrc/refex_aggregate.cu through the C-ABI in include/graphrole_b200.h).
"""
from typing import Dict, Hashable, List, Optional, Tuple

import numpy as np
import pandas as pd
import torch

from graphrole_b200.features.prune import DeviceFeaturePruner
from graphrole_b200.graph import interface
from graphrole_b200.types import DataFrameDict, DataFrameLike

# aggregations the kernel fuses; anything else has no GPU implementation and is refused
_AGG_BY_OBJECT = {
    'sum': 'sum', 'mean': 'mean',
    np.sum: 'sum', np.mean: 'mean', np.nansum: 'sum', np.nanmean: 'mean',
    pd.DataFrame.sum: 'sum', pd.DataFrame.mean: 'mean',
    pd.Series.sum: 'sum', pd.Series.mean: 'mean',
    sum: 'sum',
}


def _resolve_aggs(aggs: List) -> List[Tuple[str, str]]:
    """[(kernel op, display name)] for a user `aggs` list; display name is what pandas would
    put in the aggregated frame's index (the string itself or the callable's __name__)."""
    resolved = []
    for agg in aggs:
        try:
            op = _AGG_BY_OBJECT.get(agg)
        except TypeError:       # unhashable
            op = None
        if op is None:
            raise ValueError(
                f'aggregation {agg!r} is not supported: the sm_100a kernel implements sum and '
                f'mean (by name, numpy or pandas callables) and there is no CPU fallback')
        name = agg if isinstance(agg, str) else getattr(agg, '__name__', op)
        resolved.append((op, name))
    return resolved


class RecursiveFeatureExtractor:

    """ Compute recursive features for nodes of a graph """

    supported_graph_libs = interface.get_supported_graph_libraries()

    default_aggs = [
        pd.DataFrame.sum,
        pd.DataFrame.mean,
    ]

    # The O(n) part of pruning (binning + pairwise distances, prune.py:104-108) runs on the GPU;
    # like the aggregation it raises when there is no CUDA device -- there is no silent host
    # path.  (Unit tests of the host-side bookkeeping inject the NumPy FeaturePruner here.)
    pruner_class = DeviceFeaturePruner

    def __init__(
        self,
        G,
        max_generations: int = 10,
        aggs: Optional[List] = None,
        **kwargs
    ) -> None:
        """
        :param G: graph object from a supported graph package (or a CSRGraph)
        :param max_generations: maximum levels of recursion
        :param aggs: optional list of aggregations for each recursive generation
          (sum / mean in any order, by name or as numpy / pandas callables)
        :kwargs: kwargs accepted by the relevant graph interface
        """
        graph_class = interface.get_interface(G)
        if graph_class is None:
            raise TypeError(f'Input graph G must be from one of the following '
                            f'supported libraries: {self.supported_graph_libs}')
        graph = graph_class(G, **kwargs)
        if graph.get_num_edges() == 0:
            raise ValueError('Input graph G must contain at least one edge')

        self.graph = graph
        self.max_generations = max_generations
        self.aggs = aggs if aggs else self.default_aggs
        self._agg_ops = _resolve_aggs(self.aggs)

        self.generation_count = 0
        # binned-feature distance threshold used for pruning; tracks generation_count
        self._feature_group_thresh = 0
        # all currently retained features (rows: nodes, sorted by label)
        self._features = pd.DataFrame()
        # generation -> {feature: {node: value}} of the features retained in that generation
        self._final_features: Dict[int, DataFrameDict] = {}

    # ---- driver (host) ------------------------------------------------------------------
    def extract_features(self) -> DataFrameLike:
        """Run the recursion (or return the memoised result) and return the feature frame."""
        if self._final_features:
            return self._finalize_features()

        self._update(self.graph.get_neighborhood_features())
        for generation in range(1, self.max_generations):
            self.generation_count = generation
            self._feature_group_thresh = generation
            self._update(self._get_next_features())
            if not self._final_features[generation]:
                break
        return self._finalize_features()

    def _finalize_features(self) -> DataFrameLike:
        """All retained features; latest generation's columns first (a name retained twice
        keeps the values of its earliest generation, like ChainMap lookup does)."""
        columns: Dict[Hashable, Dict] = {}
        for generation in reversed(list(self._final_features)):
            for name in self._final_features[generation]:
                columns.setdefault(name, None)
        for generation in self._final_features:
            for name, values in self._final_features[generation].items():
                if columns[name] is None:
                    columns[name] = values
        return pd.DataFrame(columns)

    # ---- hot path A -----------------------------------------------------------------------
    def _get_next_features(self) -> DataFrameLike:
        """Next generation of candidate features: sum / mean of every node's neighbours'
        previous-generation retained features.  One kernel launch; rows come back in
        graph.get_nodes() order, columns agg-major with names `<feature>(<agg>)`."""
        prev_features = list(self._final_features[self.generation_count - 1].keys())
        csr = self.graph.to_csr()
        row_labels = list(csr.node_labels())
        prev = self._features.reindex(index=row_labels, columns=prev_features)
        sums, means = _aggregate_on_device(csr, prev.to_numpy(dtype=np.float64, na_value=np.nan))

        blocks, names = [], []
        for op, display in self._agg_ops:
            blocks.append(sums if op == 'sum' else means)
            names.extend(f'{col}({display})' for col in prev_features)
        values = np.concatenate(blocks, axis=1) if blocks else np.zeros((len(row_labels), 0))

        nodes = list(self.graph.get_nodes())
        if nodes != row_labels:
            row_of = {label: i for i, label in enumerate(row_labels)}
            values = values[[row_of[node] for node in nodes]]
        return pd.DataFrame(values, index=nodes, columns=names)

    def _update(self, features: DataFrameLike) -> None:
        """Merge a generation's candidate features, prune, and record what was retained."""
        merged = pd.concat([self._features, features], axis=1, sort=True).fillna(0)
        pruner = self.pruner_class(self._final_features, self._feature_group_thresh)
        redundant = pruner.prune_features(merged)
        self._features = merged.drop(columns=redundant)
        retained = features.columns.difference(redundant)
        self._final_features[self.generation_count] = \
            as_frame(self._features[retained]).to_dict()

    @staticmethod
    def _aggregated_df_to_dict(agg_df: DataFrameLike) -> Dict[str, float]:
        """Flatten an aggregated frame (index: aggregation, columns: features) or Series
        (name: aggregation) into {'<feature>(<agg>)': value}, aggregation-major."""
        if isinstance(agg_df, pd.Series):
            agg_df = agg_df.to_frame().T
        flat = {}
        for agg_name, row in agg_df.iterrows():
            for feature, value in row.items():
                flat[f'{feature}({agg_name})'] = value
        return flat


def _aggregate_on_device(csr, X: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """(sum, mean) over each row's out-neighbours of X [n, p] (float64 in / out, fp32 on the
    device).  NaN entries -- neighbours the frame has no row for -- are skipped the way
    pandas' skipna reductions skip them (extract.py:107-113)."""
    n, p = X.shape
    if p == 0:
        return np.zeros((n, 0)), np.zeros((n, 0))
    handle = csr.handle()
    device = handle.device
    X = np.ascontiguousarray(X)
    missing = np.isnan(X)
    if missing.any():
        # aggregate [values with NaN->0 | 1 where present]: mean = sum / count of present
        packed = np.concatenate([np.where(missing, 0.0, X), (~missing).astype(np.float64)], axis=1)
        out = handle.aggregate(torch.from_numpy(packed).to(device, torch.float32))
        out = out.double().cpu().numpy()
        sums, counts = out[:, :p], out[:, p:2 * p]
        with np.errstate(invalid='ignore', divide='ignore'):
            means = np.where(counts > 0, sums / counts, 0.0)
        return sums, means
    out = handle.aggregate(torch.from_numpy(X).to(device, torch.float32))
    out = out.double().cpu().numpy()
    return out[:, :p], out[:, p:]


def as_frame(df_like: DataFrameLike) -> pd.DataFrame:
    """pd.Series -> one-column pd.DataFrame; a DataFrame is returned unchanged."""
    return df_like.to_frame() if isinstance(df_like, pd.Series) else df_like
