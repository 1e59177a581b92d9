"""Graph plugin contract (same method names and meaning as
graphrole/graph/interface/base.py:9-83) plus the array export the CUDA path consumes."""
from abc import ABC, abstractmethod
from typing import Iterable, List

import pandas as pd

from graphrole_b200.graph.csr import CSRGraph
from graphrole_b200.types import Node


class BaseGraphInterface(ABC):
    """Adapter between a graph library object and the extractors.

    A subclass implements the five abstract methods.  `to_csr()` has a generic implementation
    on top of get_nodes()/get_neighbors() so that any third-party adapter written against the
    reference's contract feeds the GPU path unchanged; adapters that can export arrays
    directly override it.
    """

    attribute_feature_prefix = 'attribute'

    def get_neighborhood_features(self) -> pd.DataFrame:
        """Level-0 features: local columns then egonet columns, rows sorted by node label."""
        parts = [self._get_local_features(), self._get_egonet_features()]
        return pd.concat(parts, axis=1, sort=True).sort_index()

    def to_csr(self) -> CSRGraph:
        """Flatten get_nodes()/get_neighbors() into a CSRGraph (rows in sorted-label order)."""
        cached = getattr(self, '_csr_cache', None)
        if cached is None:
            cached = CSRGraph.from_neighbors(self.get_nodes(), self.get_neighbors)
            self._csr_cache = cached
        return cached

    def _set_attribute_kwargs(self, **kwargs) -> None:
        """attributes: use numeric node attributes as features; attributes_include: only these;
        attributes_exclude: never these (wins over include)."""
        self._attrs: bool = kwargs.get('attributes', False)
        self._attrs_include: List[str] = kwargs.get('attributes_include', [])
        self._attrs_exclude: List[str] = kwargs.get('attributes_exclude', [])

    @classmethod
    def _attribute_feature_name(cls, attr_name: str) -> str:
        return f'{cls.attribute_feature_prefix}_{attr_name}'

    @abstractmethod
    def get_num_edges(self) -> int:
        """Number of edges in the graph."""

    @abstractmethod
    def get_nodes(self) -> Iterable[Node]:
        """Iterable of node labels."""

    @abstractmethod
    def get_neighbors(self, node: Node) -> Iterable[Node]:
        """Iterable of the (out-)neighbours of `node`."""

    @abstractmethod
    def _get_local_features(self) -> pd.DataFrame:
        """Per-node local features (degrees, optional attributes)."""

    @abstractmethod
    def _get_egonet_features(self) -> pd.DataFrame:
        """Per-node egonet features (internal / external edge weight)."""
