"""Adapter for graphrole_b200.graph.csr.CSRGraph objects (arrays, possibly already in HBM)."""
from typing import Iterable

import pandas as pd

from graphrole_b200.graph import level0
from graphrole_b200.graph.csr import CSRGraph
from graphrole_b200.graph.interface.base import BaseGraphInterface
from graphrole_b200.types import Node


class CSRInterface(BaseGraphInterface):

    def __init__(self, G: CSRGraph, **kwargs) -> None:
        self.G = G
        self.directed = G.directed
        self._set_attribute_kwargs(**kwargs)
        self._row_of = None
        self._level0_frames = None

    def to_csr(self) -> CSRGraph:
        return self.G

    def get_num_edges(self) -> int:
        return self.G.num_edges()

    def get_nodes(self) -> Iterable[Node]:
        return self.G.node_labels()

    def get_neighbors(self, node: Node) -> Iterable[Node]:
        if self.G.labels is None:
            row = int(node)
        else:
            if self._row_of is None:
                self._row_of = {label: i for i, label in enumerate(self.G.labels)}
            row = self._row_of[node]
        lo, hi = int(self.G.rowptr[row]), int(self.G.rowptr[row + 1])
        cols = self.G.colidx[lo:hi].tolist()
        return cols if self.G.labels is None else [self.G.labels[c] for c in cols]

    # A graph whose arrays already live in HBM gets its level-0 features from the GPU kernels
    # (csrc/level0.cu); host-resident arrays use the scipy closed forms.  The choice follows
    # where the caller put the data, it is not a fallback.
    def _device_frames(self):
        if self._level0_frames is None:
            self._level0_frames = level0.device_feature_frames(self.G)
        return self._level0_frames

    def _get_local_features(self) -> pd.DataFrame:
        if self.G.rowptr.is_cuda:
            return self._device_frames()[0]
        return level0.local_degree_features(self.G)

    def _get_egonet_features(self) -> pd.DataFrame:
        if self.G.rowptr.is_cuda:
            return self._device_frames()[1]
        return level0.egonet_features(self.G)
