"""Pins oracle/refex_oracle.py (the CPU restatement of hot path A) against golden vectors
produced by the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pandas as pd
import pytest

from graphrole_b200.graph.csr import CSRGraph
from oracle import refex_oracle as oracle
from helpers import (PATH4_EXPECTED, RANDOM_CASES, frame_from_json, graph_from_json,
                     random_case)


def test_path4_known_answer(refex_cases):
    """The reference's own KAT (tests/test_features/test_extract.py:104-122)."""
    case = refex_cases['path4']
    G = graph_from_json(case['graph'])
    lvl0 = frame_from_json(case['level0'])
    got = oracle.next_features_frame(lvl0, lvl0.columns, G.nodes, lambda v: G[v].keys())
    expected = pd.DataFrame(PATH4_EXPECTED)
    np.testing.assert_allclose(got.sort_index(axis=1).sort_index(axis=0).values,
                               expected.sort_index(axis=1).sort_index(axis=0).values)
    # and the reference's actual output incl. row / column order
    ref = frame_from_json(case['next'])
    assert list(got.columns) == list(ref.columns)
    assert list(got.index) == list(ref.index)
    np.testing.assert_allclose(got.values, ref.values, rtol=1e-14)


@pytest.mark.parametrize('name', ['dangling', 'directed_weighted', 'undirected_weighted',
                                  'attributes'])
def test_small_graph_cases(refex_cases, name):
    case = refex_cases[name]
    G = graph_from_json(case['graph'], case.get('node_attrs'))
    lvl0 = frame_from_json(case['level0'])
    got = oracle.next_features_frame(lvl0, lvl0.columns, G.nodes, lambda v: G[v].keys())
    ref = frame_from_json(case['next'])
    assert list(got.columns) == list(ref.columns)
    assert list(got.index) == list(ref.index)       # get_nodes() order, not sorted
    np.testing.assert_allclose(got.values, ref.values, rtol=1e-13, atol=1e-13)
    assert not np.isnan(got.values).any()


@pytest.mark.parametrize('name', RANDOM_CASES)
def test_random_graphs_all_restatements(refex_random, name):
    directed, n, src, dst, X, out_index, out_cols, out = random_case(refex_random, name)
    g = CSRGraph.from_edges(src, dst, n=n, directed=directed)
    rp, ci = g.host_arrays()
    S, M = oracle.aggregate_csr(rp, ci, X)
    got = np.concatenate([S, M], axis=1)[out_index]
    np.testing.assert_allclose(got, out, rtol=1e-12, atol=1e-12)
    # definitional loops
    S2, M2 = oracle.aggregate_loops(rp, ci, X)
    np.testing.assert_allclose(S2, S, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(M2, M, rtol=1e-12, atol=1e-12)
    # C restatement reads float32 features: compare on float32-representable input
    X32 = X.astype(np.float32)
    S3, M3 = oracle.aggregate_rows_c(np.arange(n), rp, ci, X32)
    S4, M4 = oracle.aggregate_csr(rp, ci, X32.astype(np.float64))
    np.testing.assert_allclose(S3, S4, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(M3, M4, rtol=1e-13, atol=1e-13)


def test_pandas_chain_port_matches(refex_random):
    """The per-node pandas chain used as the CPU timing port gives the reference's numbers."""
    directed, n, src, dst, X, out_index, out_cols, out = random_case(refex_random,
                                                                     'er_undirected')
    g = CSRGraph.from_edges(src, dst, n=n, directed=directed)
    rp, ci = g.host_arrays()
    feats = pd.DataFrame(X, index=np.arange(n), columns=[f'f{j}' for j in range(X.shape[1])])
    rows = np.arange(0, n, 7)
    got = oracle.pandas_chain_rows(feats, feats.columns, rows, rp, ci)
    assert list(got.columns) == out_cols
    pos = {int(v): i for i, v in enumerate(out_index)}
    np.testing.assert_allclose(got.values, out[[pos[int(r)] for r in rows]], rtol=1e-12)


def test_karate_recursion_with_forced_retained_sets(refex_cases):
    """Replaying the reference's per-generation retained sets through the oracle reproduces
    the reference's final table and the notebook's printed table (example.ipynb cell 3)."""
    case = refex_cases['karate']
    G = graph_from_json(case['graph'])
    feats = frame_from_json(case['level0'])
    retained = case['retained_by_generation']
    final = frame_from_json(case['features'])
    prev = retained['0']
    for gen in range(1, case['generation_count'] + 1):
        nxt = oracle.next_features_frame(feats, prev, sorted(G.nodes), lambda v: G[v].keys())
        keep = retained[str(gen)]
        feats = pd.concat([feats, nxt[keep]], axis=1)
        prev = keep
    for col in final.columns:
        np.testing.assert_allclose(feats[col].values, final[col].values, rtol=1e-12)
    nb = case['notebook_table']
    table = pd.DataFrame(nb['values'], index=nb['index'], columns=nb['columns'])
    assert nb['generations'] == case['generation_count'] == 3
    assert list(table.columns) == list(final.columns)
    np.testing.assert_allclose(final[table.columns].values, table.values, atol=5.1e-7)


def test_empty_and_degenerate_inputs():
    rp = np.zeros(4, dtype=np.int64)
    S, M = oracle.aggregate_csr(rp, np.zeros(0, dtype=np.int64), np.ones((3, 2)))
    assert S.shape == (3, 2) and not S.any() and not M.any()
    S, M = oracle.aggregate_rows_c(np.arange(3), rp, np.zeros(0, dtype=np.int32),
                                   np.ones((3, 2), dtype=np.float32))
    assert not S.any() and not M.any()


def test_level_check_f64_is_the_float64_recursion():
    """oracle.level_check_f64 (C, whole level, float64 carried between levels) against
    aggregate_csr, and its error report against a perturbed 'GPU' output."""
    rng = np.random.RandomState(0)
    n, d = 4000, 6
    src, dst = rng.randint(0, 3000, 30000), rng.randint(0, n, 30000)   # rows >= 3000 are empty
    import scipy.sparse as sp
    A = sp.csr_matrix((np.ones(30000), (src, dst)), shape=(n, n))
    A.sum_duplicates()
    rp, ci = A.indptr.astype(np.int64), A.indices.astype(np.int32)
    X = rng.rand(n, d)
    for _ in range(3):
        S, M = oracle.aggregate_csr(rp, ci, X)
        gpu = np.concatenate([S, M], axis=1).astype(np.float32)
        nxt, es, em, bad = oracle.level_check_f64(rp, ci, X, gpu, threads=2)
        np.testing.assert_allclose(nxt, M, rtol=1e-15, atol=0)
        assert es < 1e-7 and em < 1e-7 and bad == 0
        X = nxt
    S, M = oracle.aggregate_csr(rp, ci, X)
    gpu = np.concatenate([S, M], axis=1).astype(np.float32)
    gpu[5, 2] *= 1.001
    gpu[3500, 1] = 1.0                                  # an empty row must stay exactly zero
    _, es, em, bad = oracle.level_check_f64(rp, ci, X, gpu)
    assert es == pytest.approx(1e-3, rel=1e-3) and bad == 1
