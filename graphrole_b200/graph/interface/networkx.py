"""Adapter for networkx Graph / DiGraph objects.

Same contract as graphrole/graph/interface/networkx.py; the graph is flattened once to CSR
(with weights) and both the level-0 features and the recursion work on the arrays.
"""
from numbers import Number
from typing import Iterable

import numpy as np
import pandas as pd

from graphrole_b200.graph import level0
from graphrole_b200.graph.csr import CSRGraph
from graphrole_b200.graph.interface.base import BaseGraphInterface
from graphrole_b200.types import Node


class NetworkxInterface(BaseGraphInterface):

    def __init__(self, G, **kwargs) -> None:
        self.G = G
        self.directed = G.is_directed()
        self._set_attribute_kwargs(**kwargs)
        self._csr_cache = None

    def get_num_edges(self) -> int:
        return self.G.number_of_edges()

    def get_nodes(self) -> Iterable[Node]:
        return self.G.nodes

    def get_neighbors(self, node: Node) -> Iterable[Node]:
        # successors for a DiGraph, neighbours otherwise; unique by construction
        return self.G[node].keys()

    def to_csr(self) -> CSRGraph:
        """One pass over G.adjacency(): rows in sorted-label order, weight default 1."""
        if self._csr_cache is not None:
            return self._csr_cache
        labels = sorted(self.G.nodes)
        row_of = {label: i for i, label in enumerate(labels)}
        src, dst, wts = [], [], []
        integral = True
        for u, nbrs in self.G.adjacency():
            ru = row_of[u]
            for v, data in nbrs.items():
                w = data.get('weight', 1)
                integral = integral and isinstance(w, (int, np.integer))
                src.append(ru)
                dst.append(row_of[v])
                wts.append(w)
        # adjacency() already lists both directions of an undirected edge: build as arcs
        csr = CSRGraph.from_edges(src, dst, n=len(labels), directed=True, weights=wts,
                                  labels=labels, weights_integral=integral)
        csr.directed = self.directed
        self._csr_cache = csr
        return csr

    def _get_local_features(self) -> pd.DataFrame:
        features = level0.local_degree_features(self.to_csr())
        if self._attrs:
            features = pd.concat([features, self._get_attribute_features()], axis=1)
        return features.fillna(0)

    def _get_egonet_features(self) -> pd.DataFrame:
        return level0.egonet_features(self.to_csr())

    def _get_attribute_features(self) -> pd.DataFrame:
        """Numeric node attributes as `attribute_<name>` columns, 0 where a node lacks one."""
        excluded = set(self._attrs_exclude)
        names = [a for a in self._attrs_include if a not in excluded]
        discover = not self._attrs_include
        columns = {self._attribute_feature_name(a): {} for a in names}
        for node, attrs in self.G.nodes(data=True):
            if discover:
                items = [(a, v) for a, v in attrs.items() if a not in excluded]
            else:
                items = [(a, attrs.get(a, 0)) for a in names]
            for attr, value in items:
                if isinstance(value, Number):
                    columns.setdefault(self._attribute_feature_name(attr), {})[node] = value
        return pd.DataFrame(columns).fillna(0)
