"""CPU tests of everything around the kernels: the C-ABI surface, graph ingest, level-0
features, pruning and the extractor / role-extractor bookkeeping (reference behaviours
restated from /root/reference/tests; values from the committed golden fixtures)."""
import ctypes
import os
import re

import networkx as nx
import numpy as np
import pandas as pd
import pytest

from graphrole_b200 import RecursiveFeatureExtractor, RoleExtractor, _native
from graphrole_b200.features.extract import _resolve_aggs, as_frame
from graphrole_b200.features.prune import FeaturePruner, vertical_log_binning
from graphrole_b200.graph import interface
from graphrole_b200.graph.csr import CSRGraph
from graphrole_b200.roles import description_length as dl
from graphrole_b200.roles import factor
from conftest import ROOT
from helpers import frame_from_json, graph_from_json


# ---- C-ABI --------------------------------------------------------------------------------

def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'graphrole_b200.h')).read()
    declared = set(re.findall(r'\b(gr_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert 'sm_100a' in _native.version()
    assert _native.launch_count() >= 0


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'graphrole_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in text.replace('# oracle', ''), os.path.join(dirpath, f)


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    G = nx.Graph([('a', 'b'), ('a', 'c'), ('c', 'd')])
    rfe = RecursiveFeatureExtractor(G, aggs=['sum', 'mean'])
    with pytest.raises(Exception):
        rfe.extract_features()
    with pytest.raises(Exception):
        factor.get_nmf_decomposition(np.random.rand(20, 30), 3)


# ---- graph ingest / plugin API -------------------------------------------------------------

def test_get_interface_dispatch():
    assert interface.get_interface(nx.Graph()) is interface.NetworkxInterface
    assert interface.get_interface(CSRGraph([0, 0], [])) is interface.CSRInterface

    class SomeGraph:
        pass
    assert interface.get_interface(SomeGraph) is None or interface.get_interface(SomeGraph())\
        is None
    assert set(interface.get_supported_graph_libraries()) >= {'networkx', 'graphrole_b200'}


def test_unknown_graph_type_and_empty_graph_raise():
    class SomeGraph:
        pass
    with pytest.raises(TypeError):
        RecursiveFeatureExtractor(SomeGraph)
    with pytest.raises(ValueError):
        RecursiveFeatureExtractor(nx.Graph())


def test_unsupported_aggregation_is_refused():
    with pytest.raises(ValueError):
        RecursiveFeatureExtractor(nx.Graph([(0, 1)]), aggs=['sum', 'max'])
    assert _resolve_aggs([np.sum, np.mean]) == [('sum', 'sum'), ('mean', 'mean')]
    assert _resolve_aggs(RecursiveFeatureExtractor.default_aggs) == [('sum', 'sum'),
                                                                     ('mean', 'mean')]
    assert _resolve_aggs(['mean']) == [('mean', 'mean')]


def test_csr_from_networkx_sorted_labels_unique_successors():
    G = nx.DiGraph([(5, 2), (2, 9), (9, 5), (1, 5), (5, 5), (2, 0)])
    itf = interface.get_interface(G)(G)
    csr = itf.to_csr()
    assert csr.labels == [0, 1, 2, 5, 9]
    rp, ci = csr.host_arrays()
    nbrs = {csr.labels[i]: [csr.labels[c] for c in ci[rp[i]:rp[i + 1]]] for i in range(csr.n)}
    assert nbrs == {0: [], 1: [5], 2: [0, 9], 5: [2, 5], 9: [5]}
    # generic plugin path gives the same arrays
    generic = CSRGraph.from_neighbors(itf.get_nodes(), itf.get_neighbors)
    np.testing.assert_array_equal(generic.host_arrays()[0], rp)
    np.testing.assert_array_equal(generic.host_arrays()[1], ci)


def test_csr_from_edges_symmetrises_and_dedupes():
    g = CSRGraph.from_edges([0, 0, 1, 2, 2], [1, 1, 2, 2, 0], n=4)
    rp, ci = g.host_arrays()
    assert rp.tolist() == [0, 2, 4, 7, 7]
    assert ci.tolist() == [1, 2, 0, 2, 0, 1, 2]
    assert g.num_edges() == 4          # 0-1, 1-2, 0-2 and the self loop
    with pytest.raises(ValueError):
        CSRGraph.from_edges([0], [9], n=3)


def test_csr_from_edges_undirected_repeats_keep_the_last_weight_both_ways():
    """Re-adding an undirected edge the other way round replaces its weight for both
    directions (what nx.Graph.add_edge does)."""
    g = CSRGraph.from_edges([0, 1, 2], [1, 0, 2], n=3, weights=[5.0, 7.0, 2.0])
    rp, ci = g.host_arrays()
    assert rp.tolist() == [0, 1, 2, 3] and ci.tolist() == [1, 0, 2]
    assert g.weights.tolist() == [7.0, 7.0, 2.0]
    G = nx.Graph()
    G.add_edge(0, 1, weight=5.0)
    G.add_edge(1, 0, weight=7.0)
    assert G[0][1]['weight'] == 7.0


def test_csr_to_keeps_every_field():
    """a weighted graph moved to another device must keep its weights (level-0
    features are weighted, networkx.py:48-83)."""
    g = CSRGraph.from_edges([0, 1, 2], [1, 2, 0], n=3, weights=[2.0, 3.5, 4.0], labels=list('abc'),
                            weights_integral=False)
    moved = g.to('meta')
    assert moved is not g and moved.weights is not None
    assert moved.weights.shape == g.weights.shape and moved.weights.device.type == 'meta'
    assert moved.labels == g.labels and moved.weights_integral is False
    assert moved.directed == g.directed and moved.n_cols == g.n_cols
    assert g.to('cpu') is g


def test_csr_save_load_and_edge_list_file(tmp_path):
    g = CSRGraph.from_edges([0, 0, 3], [1, 2, 3], n=4, weights=[1.5, 2.0, 4.0],
                            labels=['a', 'b', 'c', 'd'], weights_integral=False)
    path = str(tmp_path / 'g.npz')
    g.save(path)
    h = CSRGraph.load(path)
    assert h.labels == g.labels and h.directed == g.directed and h.n_cols == g.n_cols
    assert torch_equal(h.rowptr, g.rowptr) and torch_equal(h.colidx, g.colidx)
    assert torch_equal(h.weights, g.weights) and h.weights_integral is False
    # text edge list: integer labels sort numerically, comments are skipped
    txt = tmp_path / 'edges.txt'
    txt.write_text('# u v w\n10 2 1.0\n2 7 3.0\n10 7 2.0\n')
    e = CSRGraph.from_edge_list_file(str(txt), weighted=True)
    assert e.labels == [2, 7, 10] and e.n == 3 and e.nnz == 6
    G = nx.Graph()
    G.add_weighted_edges_from([(10, 2, 1.0), (2, 7, 3.0), (10, 7, 2.0)])
    a = interface.CSRInterface(e).get_neighborhood_features()
    b = interface.NetworkxInterface(G).get_neighborhood_features()
    pd.testing.assert_frame_equal(a.astype(float), b.astype(float))
    named = tmp_path / 'named.csv'
    named.write_text('x,y\ny,z\n')
    d = CSRGraph.from_edge_list_file(str(named), directed=True, delimiter=',')
    assert d.labels == ['x', 'y', 'z'] and d.nnz == 2 and d.directed


def torch_equal(a, b):
    import torch
    return torch.equal(a, b)


@pytest.mark.parametrize('name', ['path4', 'dangling', 'directed_weighted', 'undirected_weighted',
                                  'karate', 'karate_weighted', 'attributes', 'iface_undirected',
                                  'iface_directed_weighted'])
def test_level0_features_match_reference(refex_cases, name):
    case = refex_cases[name]
    G = graph_from_json(case['graph'], case.get('node_attrs'))
    kw = {'attributes': True} if name == 'attributes' else {}
    got = interface.get_interface(G)(G, **kw).get_neighborhood_features()
    ref = frame_from_json(case['level0'])
    assert list(got.columns) == list(ref.columns)
    assert list(got.index) == list(ref.index)
    np.testing.assert_allclose(got.values.astype(float), ref.values, rtol=1e-12)


def test_csr_interface_level0_equals_networkx_interface():
    G = nx.gnm_random_graph(60, 200, seed=0)
    a = interface.NetworkxInterface(G).get_neighborhood_features()
    csr = interface.NetworkxInterface(G).to_csr()
    b = interface.CSRInterface(csr).get_neighborhood_features()
    pd.testing.assert_frame_equal(a, b)
    itf = interface.CSRInterface(csr)
    assert sorted(itf.get_neighbors(3)) == sorted(G[3].keys())
    assert itf.get_num_edges() == G.number_of_edges()


def test_attribute_include_exclude():
    G = nx.Graph([(0, 1), (1, 2)])
    for node in G.nodes:
        G.nodes[node].update(a=float(node), b=2.0 * node, c='text')
    f = interface.NetworkxInterface(G, attributes=True, attributes_include=['b', 'a'],
                                    attributes_exclude=['a'])._get_local_features()
    assert list(f.columns) == ['degree', 'attribute_b']
    f = interface.NetworkxInterface(G, attributes=True)._get_local_features()
    assert list(f.columns) == ['degree', 'attribute_a', 'attribute_b']


# ---- pruning ---------------------------------------------------------------------------------

def test_vertical_log_binning_golden(prune_cases):
    for case in prune_cases['binning']:
        got = vertical_log_binning(np.array(case['arr']), case['frac'])
        assert got.tolist() == case['binned']
    with pytest.raises(ValueError):
        vertical_log_binning(np.arange(3), 1.0)


def test_vertical_log_binning_reference_table():
    """Spot values restated from the reference's table test (test_prune.py:18-85 style)."""
    assert vertical_log_binning(np.array([1, 1, 1, 1])).tolist() == [0, 0, 0, 0]
    assert vertical_log_binning(np.array([1, 2, 3, 4])).tolist() == [0, 0, 1, 2]
    assert vertical_log_binning(np.array([4, 3, 2, 1])).tolist() == [2, 1, 0, 0]


def test_prune_features_golden(prune_cases):
    for case in prune_cases['prune']:
        feats = pd.DataFrame(case['values'], columns=case['columns'])
        gens = {int(k): {name: {} for name in v} for k, v in case['generations'].items()}
        got = FeaturePruner(gens, case['thresh']).prune_features(feats)
        assert sorted(got) == case['dropped']


def test_pruner_on_the_references_own_tables(prune_reference_tables):
    """Host pruner against every case of the reference's known-answer tables
    (tests/test_features/test_prune.py:17-99, 119-153, 155-185)."""
    t = prune_reference_tables
    for case in t['binning']:
        for arr in (np.array(case['arr']), pd.Series(case['arr'], dtype=float)):
            assert vertical_log_binning(arr, case['frac']).tolist() == case['binned'], case['name']
    for case in t['prune']:
        feats = pd.DataFrame(case['values'], columns=case['columns'])
        gens = {int(k): {name: {} for name in v} for k, v in case['generations'].items()}
        got = FeaturePruner(gens, case['thresh']).prune_features(feats)
        assert sorted(got) == case['dropped']
    for case in t['group']:
        feats = pd.DataFrame(case['values'], columns=case['columns'])
        groups = FeaturePruner({0: {'b': {}, 'a': {}}, 1: {'c': {}, 'd': {}}},
                               case['thresh'])._group_features(feats)
        assert sorted(sorted(g) for g in groups) == case['groups']


def test_oldest_feature_tie_break():
    pruner = FeaturePruner({0: {'b': {}, 'a': {}}, 1: {'c': {}}}, 0)
    assert pruner._get_oldest_feature({'c', 'b', 'a'}) == 'a'
    assert pruner._get_oldest_feature({'c'}) == 'c'
    assert pruner._get_oldest_feature({'z', 'y'}) == 'y'


# ---- extractor bookkeeping (no GPU needed: state is seeded like the reference's tests) -------

def _seeded_rfe():
    G = nx.Graph([('a', 'b'), ('a', 'c'), ('c', 'd')])
    rfe = RecursiveFeatureExtractor(G, aggs=[np.sum, np.mean])
    rfe._features = rfe.graph.get_neighborhood_features()
    rfe._final_features = {0: rfe._features.to_dict()}
    rfe.generation_count = 1
    return rfe


def test_update_prunes_and_records_retained():
    """Behaviour of tests/test_features/test_extract.py:124-159."""
    rfe = _seeded_rfe()
    rfe.pruner_class = FeaturePruner     # bookkeeping under test; the GPU pruner has its own tests
    existing = rfe._features
    rng = np.random.RandomState(0)
    new = pd.concat([
        pd.DataFrame(existing['degree'].values, columns=['degree2'], index=existing.index),
        pd.DataFrame(rng.randn(existing.shape[0], 2), columns=['a', 'b'], index=existing.index),
    ], axis=1)
    rfe._update(new)
    expected = pd.concat([existing[['degree', 'external_edges']], new[['a', 'b']]], axis=1)
    pd.testing.assert_frame_equal(rfe._features, expected)
    final = rfe._finalize_features()
    expected_final = pd.concat([existing, new[['a', 'b']]], axis=1)
    pd.testing.assert_frame_equal(final.sort_index(axis=1), expected_final.sort_index(axis=1))


def test_aggregated_df_to_dict():
    index = ['sum', 'mean']
    columns = ['feature1', 'feature2', 'feature3']
    df = pd.DataFrame(np.arange(6).reshape(2, 3), columns=columns, index=index)
    got = RecursiveFeatureExtractor._aggregated_df_to_dict(df)
    assert got == {'feature1(sum)': 0, 'feature2(sum)': 1, 'feature3(sum)': 2,
                   'feature1(mean)': 3, 'feature2(mean)': 4, 'feature3(mean)': 5}
    assert list(got) == ['feature1(sum)', 'feature2(sum)', 'feature3(sum)',
                         'feature1(mean)', 'feature2(mean)', 'feature3(mean)']
    series = pd.Series([6, 7, 8], index=columns, name='prod')
    assert RecursiveFeatureExtractor._aggregated_df_to_dict(series) == {
        'feature1(prod)': 6, 'feature2(prod)': 7, 'feature3(prod)': 8}


def test_finalize_features_latest_generation_first():
    rfe = _seeded_rfe()
    data = {'node1': {'a': 0, 'b': 1, 'c': 2, 'd': 3, 'e': 4},
            'node2': {'a': 5, 'b': 6, 'c': 7, 'd': 8, 'e': 9}}
    expected = pd.DataFrame.from_dict(data, orient='index')
    rfe._final_features = {0: expected[['a', 'b']].to_dict(), 1: expected[['c', 'd']].to_dict(),
                           2: expected['e'].to_frame().to_dict()}
    final = rfe._finalize_features()
    pd.testing.assert_frame_equal(final.sort_index(axis=1), expected.sort_index(axis=1))
    assert list(final.columns) == ['e', 'c', 'd', 'a', 'b']
    # memoisation: extract_features returns stored state without touching the graph
    pd.testing.assert_frame_equal(rfe.extract_features(), final)


def test_as_frame():
    s = pd.Series(np.arange(4.0))
    pd.testing.assert_frame_equal(as_frame(s), pd.DataFrame(s))
    f = pd.DataFrame(np.arange(4.0))
    assert as_frame(f) is f


# ---- roles: host-side pieces ---------------------------------------------------------------

def test_encode_refuses_more_bins_than_entries_before_touching_the_device():
    """The ValueError the grid search relies on (roles/extract.py:127-129) has its own class, so
    the grid does not swallow unrelated argument errors; the costs and the quantiser themselves
    run on the GPU (tests/test_rolx_gpu.py)."""
    with pytest.raises(_native.TooManyBinsError, match='should be >= n_clusters'):
        factor.encode(np.random.rand(2, 2), 16)
    assert issubclass(_native.TooManyBinsError, ValueError)
    assert dl.encoding_cost_from_counts(4, 3, 12) == 2 * 12


def test_rescale_costs_and_role_extractor_surface():
    costs = np.array([[np.nan, np.nan], [3.0, 4.0], [np.nan, 2.0]])
    got = RoleExtractor._rescale_costs(costs)
    np.testing.assert_allclose(got[1], [0.6, 0.8])
    np.testing.assert_allclose(got[2, 1], 1.0)
    rx = RoleExtractor(n_roles=None, n_role_range=(2, 5), n_bit_range=(1, 4))
    assert (rx.min_roles, rx.max_roles, rx.min_bits, rx.max_bits) == (2, 5, 1, 4)
    assert rx.roles is None and rx.role_percentage is None
    with pytest.raises(NotImplementedError):
        rx.explain()
    rx.node_role_factor = pd.DataFrame([[1.0, 3.0], [2.0, 2.0]], index=['x', 'y'],
                                       columns=['role_0', 'role_1'])
    assert rx.roles == {'x': 'role_1', 'y': 'role_0'}
    np.testing.assert_allclose(rx.role_percentage.values, [[0.25, 0.75], [0.5, 0.5]])


def test_orthonormal_basis_cholesky_qr2_and_householder_fallback():
    """The range finder's basis (factor.orthonormal_basis): CholeskyQR2 for tall sketches gives an
    orthonormal basis of the same range as Householder QR; a rank-deficient sketch (the float64
    Cholesky fails) and a short one take Householder QR itself."""
    import torch
    gen = torch.Generator().manual_seed(0)
    for m, k, cond, via_cholesky in [(20000, 18, 1e2, True), (20000, 12, 1e4, True),
                                     (20000, 12, 1e12, False), (100, 18, 10, False)]:
        U, _ = torch.linalg.qr(torch.randn(m, k, dtype=torch.float64, generator=gen))
        V, _ = torch.linalg.qr(torch.randn(k, k, dtype=torch.float64, generator=gen))
        sv = torch.logspace(0, -np.log10(cond), k, dtype=torch.float64)
        Y = ((U * sv) @ V.T).float()
        Q = factor.orthonormal_basis(Y)
        assert Q.shape == (m, k) and Q.dtype == Y.dtype
        Qd = Q.double()
        assert float((Qd.T @ Qd - torch.eye(k, dtype=torch.float64)).abs().max()) < 1e-6
        Qh = torch.linalg.qr(Y)[0]
        if not via_cholesky:
            assert torch.equal(Q, Qh)
        elif cond <= 1e2:           # every direction resolved in float32: same range
            assert float((Qd @ (Qd.T @ Qh.double()) - Qh.double()).abs().max()) < 1e-4
    # 18 columns of rank 5 (up to float32 rounding): whichever route it takes, the basis is
    # orthonormal and holds the 5-dimensional range
    A5 = torch.randn(5000, 5, generator=gen)
    Q = factor.orthonormal_basis(A5 @ torch.randn(5, 18, generator=gen)).double()
    assert float((Q.T @ Q - torch.eye(18, dtype=torch.float64)).abs().max()) < 1e-6
    B5 = torch.linalg.qr(A5.double())[0]
    assert float((Q @ (Q.T @ B5) - B5).abs().max()) < 1e-4


def test_nndsvda_init_on_a_tall_matrix_matches_sklearn_cpu():
    """roles.factor.nndsvda_init is device-agnostic torch code: on CPU tensors a tall matrix takes
    the CholeskyQR2 sketches (6000 rows >= 64 x 14 columns) and must give sklearn's start for the
    same NumPy random stream (sklearn/decomposition/_nmf.py:214-366)."""
    import torch
    sk = pytest.importorskip('sklearn.decomposition._nmf')
    rng = np.random.RandomState(1)
    X = (rng.rand(6000, 5) ** 2 * np.array([9.0, 5.0, 3.0, 2.0, 1.0])) @ rng.rand(5, 40) + \
        0.01 * rng.rand(6000, 40)
    np.random.seed(7)
    W_ref, H_ref = sk._initialize_nmf(X, 4, init='nndsvda')
    np.random.seed(7)
    W, H = factor.nndsvda_init(torch.from_numpy(X), 4)
    np.testing.assert_allclose(W.numpy(), W_ref, rtol=5e-4, atol=2e-5)
    np.testing.assert_allclose(H.numpy(), H_ref, rtol=5e-4, atol=2e-5)
