"""Adapter for python-igraph Graph objects (optional dependency; mirrors the contract of
graphrole/graph/interface/igraph.py).  Vertices are addressed by index, as in the reference."""
from typing import Iterable

import numpy as np
import pandas as pd

from graphrole_b200.graph import level0
from graphrole_b200.graph.csr import CSRGraph
from graphrole_b200.graph.interface.base import BaseGraphInterface
from graphrole_b200.types import Node


class IgraphInterface(BaseGraphInterface):

    def __init__(self, G, **kwargs) -> None:
        self.G = G
        self.directed = G.is_directed()
        self._set_attribute_kwargs(**kwargs)
        self._csr_cache = None

    def get_num_edges(self) -> int:
        return self.G.ecount()

    def get_nodes(self) -> Iterable[Node]:
        return self.G.vs().indices

    def get_neighbors(self, node: Node) -> Iterable[Node]:
        return self.G.neighbors(node, mode='out')

    def to_csr(self) -> CSRGraph:
        if self._csr_cache is None:
            edges = np.asarray(self.G.get_edgelist(), dtype=np.int64).reshape(-1, 2)
            weights = self.G.es['weight'] if 'weight' in self.G.es.attributes() else None
            integral = weights is None or all(isinstance(w, (int, np.integer)) for w in weights)
            self._csr_cache = CSRGraph.from_edges(
                edges[:, 0], edges[:, 1], n=self.G.vcount(), directed=self.directed,
                weights=weights, weights_integral=integral)
        return self._csr_cache

    def _get_local_features(self) -> pd.DataFrame:
        features = level0.local_degree_features(self.to_csr())
        if self._attrs:
            excluded = set(self._attrs_exclude)
            names = [a for a in (self._attrs_include or self.G.vs.attributes())
                     if a not in excluded]
            cols = {}
            for name in names:
                vals = self.G.vs[name] if name in self.G.vs.attributes() else [0] * len(features)
                if all(isinstance(v, (int, float, np.number)) or v is None for v in vals):
                    cols[self._attribute_feature_name(name)] = [0 if v is None else v
                                                                for v in vals]
            features = pd.concat([features, pd.DataFrame(cols, index=features.index)], axis=1)
        return features.fillna(0)

    def _get_egonet_features(self) -> pd.DataFrame:
        return level0.egonet_features(self.to_csr())
