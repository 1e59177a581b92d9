"""TEST / BENCH INFRASTRUCTURE -- drives the UNMODIFIED reference's own hot path A on the CPU.

`__graft_entry__.build()` stages the reference package (pure Python) from /root/reference into the
git-ignored oracle/_ref/ (it travels to the GPU box like the built .so files; nothing of it is in
the repository's history).  This module imports it from there and calls the reference's real
`RecursiveFeatureExtractor._get_next_features` (graphrole/features/extract.py:98-119) through the
reference's own plugin API: a `BaseGraphInterface` subclass over CSR arrays whose `get_nodes()`
yields a node SAMPLE -- the full loop would take ~145 h per level at
10 M nodes -- with the previous generation's features injected through the seam the reference's
tests use (tests/test_features/test_extract.py:87-92).

pandas here is 3.x (the reference pins < 2): its default aggs raise TypeError, so the extractor is
built with aggs=['sum', 'mean'] like every golden vector of this repository.
Only bench.py (`--impl reference`, `cpu_baseline`) and tests/ may import this module.
"""
import importlib
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, '_ref')


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, 'graphrole', '__init__.py'))


def load_reference():
    """Import the staged reference package (oracle/_ref/graphrole)."""
    if not available():
        raise ImportError(f'{REF_DIR}/graphrole is not staged: run __graft_entry__.build() in the '
                          f'container that has /root/reference')
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    return importlib.import_module('graphrole')


class SampledCsrGraph:
    """What the reference's extractor is handed: CSR arrays plus the node sample to iterate."""

    def __init__(self, rowptr, colidx, sample):
        self.rowptr, self.colidx = rowptr, colidx
        self.sample = np.asarray(sample, dtype=np.int64)


def _interface_class():
    load_reference()
    base = importlib.import_module('graphrole.graph.interface.base')

    class CsrSampleInterface(base.BaseGraphInterface):
        """BaseGraphInterface (graphrole/graph/interface/base.py:9-83) over CSR arrays; node
        labels are row numbers.  Only what path A consumes is implemented."""

        def __init__(self, G, **kwargs):
            self.G = G
            self._set_attribute_kwargs(**kwargs)

        def get_num_edges(self):
            return int(self.G.colidx.shape[0])

        def get_nodes(self):
            return [int(i) for i in self.G.sample]

        def get_neighbors(self, node):
            rp, ci = self.G.rowptr, self.G.colidx
            return ci[rp[node]:rp[node + 1]]

        def _get_local_features(self):
            raise NotImplementedError('level-0 features are not on the timed path')

        def _get_egonet_features(self):
            raise NotImplementedError('level-0 features are not on the timed path')

    return CsrSampleInterface


class ReferenceLevel:
    """The reference's extractor seeded with a feature matrix, ready to run
    `_get_next_features` on node samples of a CSR graph."""

    def __init__(self, rowptr, colidx, X):
        import pandas as pd
        graphrole = load_reference()
        registry = importlib.import_module('graphrole.graph.interface')
        registry.INTERFACES[SampledCsrGraph.__module__.split('.')[0]] = _interface_class()
        self.graph = SampledCsrGraph(np.asarray(rowptr), np.asarray(colidx), [0])
        self.rfe = graphrole.RecursiveFeatureExtractor(self.graph, aggs=['sum', 'mean'])
        X = np.asarray(X, dtype=np.float64)
        cols = [f'f{j}' for j in range(X.shape[1])]
        self.rfe._features = pd.DataFrame(X, columns=cols)
        self.rfe._final_features = {0: {c: {} for c in cols}}
        self.rfe.generation_count = 1
        self.d = X.shape[1]

    def rows(self, sample):
        """The reference's `_get_next_features` restricted to `sample` (a DataFrame)."""
        self.graph.sample = np.asarray(sample, dtype=np.int64)
        return self.rfe._get_next_features()

    def timed(self, sample):
        """(arc*features per second, arcs, seconds) of one `_get_next_features` call."""
        rp = self.graph.rowptr
        sample = np.asarray(sample, dtype=np.int64)
        t0 = time.perf_counter()
        self.rows(sample)
        dt = time.perf_counter() - t0
        arcs = int((rp[sample + 1] - rp[sample]).sum())
        return arcs * self.d / dt, arcs, dt
