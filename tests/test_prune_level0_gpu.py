"""GPU parity tests of the "next" rows of the path: vertical log binning,
pairwise Chebyshev gaps and level-0 egonet features, each through the C-ABI and checked against
oracle/prune_oracle.py (pinned to the reference's vectors in test_oracle_prune.py) and against
the committed golden fixtures.  Integer results: bit-exact."""
import networkx as nx
import numpy as np
import pandas as pd
import pytest
import torch

from graphrole_b200 import RecursiveFeatureExtractor, _native
from graphrole_b200.features.prune import DeviceFeaturePruner, FeaturePruner, vertical_log_binning
from graphrole_b200.graph import interface, level0
from graphrole_b200.graph.csr import CSRGraph
from graphrole_b200.graph.generators import barabasi_albert_csr
from oracle import prune_oracle
from helpers import frame_from_json, graph_from_json

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _bins(X, frac=0.5):
    X = torch.as_tensor(X, device=DEV)
    if X.dim() == 1:
        X = X[:, None]
    p = _native.Pruner(X.shape[0], DEV)
    try:
        return p.bin_columns(X.contiguous(), frac).cpu().numpy()
    finally:
        p.close()


# ---- binning -----------------------------------------------------------------------------------

@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_binning_matches_reference_vectors(prune_cases, dtype):
    launches = _native.launch_count()
    for case in prune_cases['binning']:
        arr = np.array(case['arr'], dtype=np.float64)
        if dtype == np.float32 and not np.array_equal(arr.astype(np.float32).astype(np.float64), arr):
            arr = arr.astype(np.float32)      # fp32 path: compare on the fp32-rounded values
            want = prune_oracle.vertical_log_binning(arr, case['frac']).tolist()
        else:
            want = case['binned']
        got = _bins(arr.astype(dtype), case['frac'])[0]
        assert got.tolist() == want
    assert _native.launch_count() > launches


def test_device_pruner_on_the_references_own_tables(prune_reference_tables):
    """GPU binning / grouping against the reference's known-answer tables
    (tests/test_features/test_prune.py:17-99, 119-153, 155-185); the empty vector never reaches
    the device (a feature matrix has at least one row)."""
    t = prune_reference_tables
    for case in t['binning']:
        if not case['arr']:
            continue
        got = _bins(np.array(case['arr'], dtype=np.float64), case['frac'])[0]
        assert got.tolist() == case['binned'], case['name']
    for case in t['prune']:
        feats = pd.DataFrame(case['values'], columns=case['columns'])
        gens = {int(k): {name: {} for name in v} for k, v in case['generations'].items()}
        got = DeviceFeaturePruner(gens, case['thresh']).prune_features(feats)
        assert sorted(got) == case['dropped']
    for case in t['group']:
        feats = pd.DataFrame(case['values'], columns=case['columns'])
        groups = DeviceFeaturePruner({0: {'b': {}, 'a': {}}, 1: {'c': {}, 'd': {}}},
                                     case['thresh'])._group_features(feats)
        assert sorted(sorted(g) for g in groups) == case['groups']


@pytest.mark.parametrize('n', [1, 2, 3, 31, 32, 33, 1000, 4097])
@pytest.mark.parametrize('frac', [0.5, 0.3, 0.9, 0.05])
def test_binning_small_sizes_heavy_ties(n, frac):
    rng = np.random.RandomState(n)
    cols = [rng.randint(0, 4, n), rng.randint(0, max(2, n // 3), n), rng.rand(n),
            np.zeros(n), np.arange(n)[::-1], rng.choice([-0.0, 0.0, 1.5, -2.0], n)]
    X = np.stack([np.asarray(c, dtype=np.float64) for c in cols], axis=1)
    got = _bins(X, frac)
    for c in range(X.shape[1]):
        want = prune_oracle.vertical_log_binning(X[:, c], frac)
        assert got[c].tolist() == want.tolist(), (n, frac, c)


@pytest.mark.parametrize('d', [1, 7, 32, 33, 70])
def test_binning_many_columns_strided_input(d):
    n = 20_011
    g = torch.Generator(device=DEV).manual_seed(d)
    big = torch.rand(n, d + 5, device=DEV, generator=g)
    big[:, ::3] = torch.floor(big[:, ::3] * 6)            # tied columns
    X = big[:, 2:2 + d]                                    # row stride d + 5, offset 2
    p = _native.Pruner(n, DEV)
    got = p.bin_columns(X).cpu().numpy()
    p.close()
    Xh = X.cpu().numpy()
    for c in range(d):
        # host restatement used by the product (pinned to the same vectors); the definitional
        # oracle is quadratic-ish in Python, so spot-check a few columns with it
        assert np.array_equal(got[c], vertical_log_binning(Xh[:, c])), c
    for c in sorted({0, d // 2, d - 1}):
        assert np.array_equal(got[c], prune_oracle.vertical_log_binning(Xh[:, c]))


def test_binning_300k_rows_equals_the_oracle_exactly():
    """A size the definitional oracle still finishes in seconds: every bin of every row."""
    n = 300_000
    rng = np.random.RandomState(9)
    g = barabasi_albert_csr(n, 6, seed=9, device=DEV)
    deg = g.out_degree().double().cpu().numpy()              # power-law ties
    X = np.stack([rng.rand(n), rng.randint(0, 7, n).astype(float), deg,
                  np.round(rng.randn(n), 1)], axis=1)
    for dtype in (np.float64, np.float32):
        Xd = X.astype(dtype)
        got = _bins(Xd)
        for c in range(X.shape[1]):
            want = prune_oracle.vertical_log_binning(Xd[:, c].astype(np.float64))
            assert np.array_equal(got[c], want), (dtype, c)


def test_binning_large_column_properties():
    """2 M rows: bins are monotone in the value, ties share a bin, bin sizes follow the
    halving rule."""
    n = 2_000_000
    g = torch.Generator(device=DEV).manual_seed(1)
    # column 0: all values distinct (a permutation); column 1: 50 distinct values
    X = torch.stack([torch.randperm(n, device=DEV, generator=g).float() * 0.5,
                     torch.randint(0, 50, (n,), device=DEV, generator=g).float()], dim=1)
    p = _native.Pruner(n, DEV)
    bins = p.bin_columns(X)
    p.close()
    for c in range(2):
        v, b = X[:, c], bins[c].long()
        order = torch.argsort(v, stable=True)
        bs = b[order]
        assert bool((bs[1:] >= bs[:-1]).all())                       # monotone in the value
        vs = v[order]
        same = vs[1:] == vs[:-1]
        assert bool((bs[1:][same] == bs[:-1][same]).all())           # ties never split
        counts = torch.bincount(b)
        assert int(counts.sum()) == n and int(counts.min()) >= 1
    counts = torch.bincount(bins[0].long()).cpu().numpy()             # distinct values: exact rule
    left = n
    for cnt in counts:
        assert cnt == max(int(0.5 * left), 1)
        left -= cnt
    assert left == 0


def test_binning_rejects_bad_arguments():
    p = _native.Pruner(10, DEV)
    X = torch.rand(10, 3, device=DEV)
    for frac in (0.0, 1.0, -0.1):
        with pytest.raises(ValueError, match='frac'):
            p.bin_columns(X, frac)
    with pytest.raises(ValueError):
        p.bin_columns(torch.rand(11, 3, device=DEV))
    with pytest.raises(ValueError):
        p.bin_columns(X.t().contiguous().t())                        # column stride != 1
    assert p.bin_columns(torch.rand(10, 0, device=DEV)).shape == (0, 10)
    p.close()


# ---- pairwise gaps ---------------------------------------------------------------------------------

@pytest.mark.parametrize('n,d', [(1, 2), (5, 3), (100, 4), (1000, 5), (777, 9), (5000, 33),
                                 (3001, 64), (3001, 95), (3001, 96), (777, 97), (2000, 100), (3001, 128),
                                 (600, 260), (100, 1024)])
def test_pairwise_gaps_match_numpy(n, d):
    rng = np.random.RandomState(d)
    bins = rng.randint(0, 12, (d, n)).astype(np.int32)
    bins[d // 2] = bins[0]                                           # one identical pair
    if d > 2:
        bins[d - 1] = bins[1] + (rng.rand(n) < 0.01)                 # one pair at distance <= 1
    p = _native.Pruner(n, DEV)
    got = p.pairwise_gaps(torch.from_numpy(bins).to(DEV)).cpu().numpy()
    p.close()
    want = np.abs(bins[:, None, :].astype(np.int64) - bins[None, :, :]).max(axis=2)
    assert np.array_equal(got, want)
    if d <= 40:
        assert np.array_equal(got, prune_oracle.chebyshev_gaps(bins))


def test_pairwise_gaps_tile_sizes_agree(monkeypatch):
    """From 96 columns on the kernel keeps 8 x 8 pair tiles in registers, below that 4 x 4
    (GR_PRUNE_TILE4=1 forces 4 x 4): same integers, many rows, ragged column count."""
    n, d = 400_003, 100
    g = torch.Generator(device=DEV).manual_seed(9)
    bins = torch.randint(0, 25, (d, n), device=DEV, generator=g, dtype=torch.int32)
    bins[98] = bins[3]
    bins[99, n - 1] += 5000
    p = _native.Pruner(n, DEV)
    wide = p.pairwise_gaps(bins)
    monkeypatch.setenv('GR_PRUNE_TILE4', '1')
    narrow = p.pairwise_gaps(bins)
    p.close()
    assert torch.equal(wide, narrow)
    assert int(wide[98, 3]) == 0 and int(wide[99].max()) >= 4976
    sub = bins[:, :5000]
    want = (sub[:, None, :].long() - sub[None, :, :].long()).abs().amax(dim=2)
    p = _native.Pruner(5000, DEV)
    assert torch.equal(p.pairwise_gaps(sub.contiguous()).long(), want)
    p.close()


def test_pairwise_gaps_large_values_and_many_rows():
    n, d = 1_000_003, 6
    g = torch.Generator(device=DEV).manual_seed(3)
    bins = torch.randint(0, 30, (d, n), device=DEV, generator=g, dtype=torch.int32)
    bins[4] = bins[2]
    bins[4, n - 1] += 100_000                                        # a single far-away row at the end
    p = _native.Pruner(n, DEV)
    got = p.pairwise_gaps(bins)
    p.close()
    want = (bins[:, None, :].long() - bins[None, :, :].long()).abs().amax(dim=2)
    assert torch.equal(got.long(), want)
    assert int(got[4, 2]) == 100_000


# ---- the pruner as the extractor uses it --------------------------------------------------------

def test_device_pruner_matches_reference_vectors(prune_cases):
    for case in prune_cases['prune']:
        feats = pd.DataFrame(case['values'], columns=case['columns'])
        gens = {int(k): {name: {} for name in v} for k, v in case['generations'].items()}
        got = DeviceFeaturePruner(gens, case['thresh']).prune_features(feats)
        assert sorted(got) == case['dropped']
        want = prune_oracle.dropped_features(case['values'], case['columns'],
                                             {int(k): v for k, v in case['generations'].items()},
                                             case['thresh'])
        assert sorted(got) == want


def test_device_pruner_equals_host_pruner_on_extractor_frames():
    G = nx.gnm_random_graph(300, 1500, seed=4)
    rfe = RecursiveFeatureExtractor(G, aggs=['sum', 'mean'])
    rfe._update(rfe.graph.get_neighborhood_features())
    rfe.generation_count = rfe._feature_group_thresh = 1
    cand = rfe._get_next_features()
    merged = pd.concat([rfe._features, cand], axis=1, sort=True).fillna(0)
    for thresh in (0, 1, 2):
        a = FeaturePruner(rfe._final_features, thresh).prune_features(merged)
        b = DeviceFeaturePruner(rfe._final_features, thresh).prune_features(merged)
        assert sorted(a) == sorted(b)


# ---- level-0 features -------------------------------------------------------------------------------

@pytest.mark.parametrize('name', ['path4', 'dangling', 'directed_weighted', 'undirected_weighted',
                                  'karate', 'karate_weighted', 'iface_undirected',
                                  'iface_directed_weighted'])
def test_level0_device_matches_reference_tables(refex_cases, name):
    case = refex_cases[name]
    G = graph_from_json(case['graph'])
    ref = frame_from_json(case['level0'])
    cols = level0.device_features(_weighted_csr(G), DEV)
    assert list(cols) == list(ref.columns)
    for col in ref.columns:
        np.testing.assert_allclose(cols[col].cpu().numpy(), ref[col].values, rtol=1e-12,
                                   err_msg=col)


def _weighted_csr(G):
    labels = sorted(G.nodes)
    row = {v: i for i, v in enumerate(labels)}
    src, dst, w = [], [], []
    for u, v, data in G.edges(data=True):
        src.append(row[u]); dst.append(row[v]); w.append(data.get('weight', 1))
    return CSRGraph.from_edges(src, dst, n=len(labels), directed=G.is_directed(), weights=w,
                               labels=labels)


@pytest.mark.parametrize('directed', [False, True])
@pytest.mark.parametrize('weighted', [False, True])
def test_level0_device_matches_oracle_random(directed, weighted):
    rng = np.random.RandomState(7 + directed + 2 * weighted)
    n, m = 400, 3000
    src, dst = rng.randint(0, n, m), rng.randint(0, n, m)          # self loops and repeats included
    w = rng.randint(1, 6, m).astype(float) + (0.25 if weighted else 0.0) * rng.randint(0, 4, m)
    g = CSRGraph.from_edges(src, dst, n=n, directed=directed, weights=w if weighted else None)
    rp, ci = g.host_arrays()
    wh = np.ones(g.nnz) if g.weights is None else g.weights.numpy()
    arcs = [(i, int(ci[k]), float(wh[k])) for i in range(n) for k in range(rp[i], rp[i + 1])]
    want = prune_oracle.level0_features(n, arcs, directed)
    got = level0.device_features(g, DEV)
    assert list(got) == list(want)
    for col in want:
        np.testing.assert_allclose(got[col].cpu().numpy(), want[col], rtol=1e-12, err_msg=col)
    # and the host closed forms the CSR adapter uses for host-resident graphs
    a = pd.concat([level0.local_degree_features(g), level0.egonet_features(g)], axis=1)
    for col in want:
        np.testing.assert_allclose(got[col].cpu().numpy(), a[col].values.astype(float), rtol=1e-12)


def test_level0_device_hub_rows_and_frames():
    """BA graph: rows above the 1024-arc cut take the CTA-per-row kernel; integers are exact."""
    g = barabasi_albert_csr(200_000, 12, seed=3, device=DEV)
    deg = g.out_degree()
    assert int(deg.max()) >= 1024
    cols = level0.device_features(g)
    assert torch.equal(cols['degree'], deg.double())
    # triangle identity on an undirected simple graph: internal = deg + triangles,
    # external = sum of neighbour degrees - deg - 2 * triangles
    rp, ci = g.host_arrays()
    import scipy.sparse as sp
    A = sp.csr_matrix((np.ones(g.nnz), ci, rp), shape=(g.n, g.n))
    tri = np.asarray((A @ A).multiply(A).sum(axis=1)).ravel() / 2
    d = np.diff(rp).astype(float)
    assert np.array_equal(cols['internal_edges'].cpu().numpy(), d + tri)
    assert np.array_equal(cols['external_edges'].cpu().numpy(), A @ d - d - 2 * tri)
    # the CSR adapter serves device-resident graphs from the kernels
    launches = _native.launch_count()
    frame = interface.CSRInterface(g).get_neighborhood_features()
    assert _native.launch_count() > launches
    assert list(frame.columns) == ['degree', 'internal_edges', 'external_edges']
    assert frame['degree'].dtype == np.int64
    assert np.array_equal(frame['internal_edges'].values, (d + tri).astype(np.int64))


def test_level0_triangle_fast_path_equals_general_kernels(refex_cases, monkeypatch):
    """Undirected + unweighted + no self loop graphs take the triangle-counting path on the
    degree-oriented graph (csrc/level0.cu); it must give exactly what the general
    sorted-intersection kernels give (GR_LEVEL0_GENERAL=1), on the reference's own unweighted
    tables and on random graphs with isolated nodes, and fall back when there is a self loop."""
    from graphrole_b200.graph.generators import erdos_renyi_csr
    for name in ('path4', 'dangling', 'karate', 'iface_undirected'):
        case = refex_cases[name]
        G = graph_from_json(case['graph'])
        labels = sorted(G.nodes)
        row = {v: i for i, v in enumerate(labels)}
        g = CSRGraph.from_edges([row[u] for u, _ in G.edges()], [row[v] for _, v in G.edges()],
                                n=len(labels), labels=labels)
        ref = frame_from_json(case['level0'])
        cols = level0.device_features(g, DEV)
        for col in ref.columns:
            np.testing.assert_array_equal(cols[col].cpu().numpy(), ref[col].values, err_msg=col)
    # the intersection has three forms (shared-memory hash table of N+(u) for rows of up to 40
    # oriented arcs, binary search over the shorter list beyond that, and -- hashed row, much
    # longer neighbour list -- binary search in the neighbour's list): a dense graph (oriented rows
    # of ~100 arcs) and a clique with a sparse periphery (short rows whose neighbours are clique
    # members with ~100 oriented arcs) reach the two that sparse graphs do not
    rng = np.random.RandomState(5)
    clique = 150
    cu, cv = np.triu_indices(clique, k=1)
    per = np.arange(clique, clique + 6000)
    pu = np.repeat(per, 3)
    pv = rng.randint(0, clique, size=pu.size)
    qu = per
    qv = rng.permutation(per)
    keep = qu != qv
    mixed = CSRGraph.from_edges(np.concatenate([cu, pu, qu[keep]]),
                                np.concatenate([cv, pv, qv[keep]]), n=clique + 6000)
    graphs = [erdos_renyi_csr(50_000, 400_000, seed=1, device=DEV),
              barabasi_albert_csr(60_000, 9, seed=2, device=DEV),
              erdos_renyi_csr(3_000, 300_000, seed=3, device=DEV),
              mixed,
              CSRGraph.from_edges([0, 1, 5], [1, 2, 6], n=9)]        # isolated nodes 3, 4, 7, 8
    for g in graphs:
        fast = level0.device_features(g, DEV)
        monkeypatch.setenv('GR_LEVEL0_GENERAL', '1')
        general = level0.device_features(g, DEV)
        monkeypatch.delenv('GR_LEVEL0_GENERAL')
        for col in general:
            assert torch.equal(fast[col], general[col]), col
    loop = CSRGraph.from_edges([0, 1, 2, 2], [1, 2, 0, 2], n=3)          # triangle + self loop
    got = level0.device_features(loop, DEV)
    want = pd.concat([level0.local_degree_features(loop), level0.egonet_features(loop)], axis=1)
    for col in got:
        np.testing.assert_array_equal(got[col].cpu().numpy(), want[col].values.astype(float))


# ---- device-resident recursion -------------------------------------------------------------------

@pytest.mark.parametrize('name', ['dangling', 'directed_weighted', 'undirected_weighted',
                                  'karate', 'karate_weighted'])
def test_device_resident_extractor_matches_reference_frames(refex_cases, name):
    """Whole extract_features() with every O(n) step on the GPU (level 0, aggregation, binning,
    pairwise gaps) against the frames the unmodified reference produced."""
    from graphrole_b200.features.device import DeviceRecursiveFeatureExtractor
    case = refex_cases[name]
    G = graph_from_json(case['graph'])
    rfe = DeviceRecursiveFeatureExtractor(_weighted_csr(G), device=DEV)
    got = rfe.extract_features()
    ref = frame_from_json(case['features'])
    assert rfe.generation_count == case['generation_count']
    assert list(got.columns) == list(ref.columns)
    assert list(got.index) == list(ref.index)
    np.testing.assert_allclose(got.values.astype(float), ref.values, rtol=1e-5, atol=1e-12)
    if name == 'karate':
        assert {str(k): sorted(v) for k, v in rfe._final_features.items()} == \
            case['retained_by_generation']
    # and identical (names, generation by generation) to the pandas-facing extractor
    host = RecursiveFeatureExtractor(G)
    host.extract_features()
    assert {k: sorted(v) for k, v in host._final_features.items()} == \
        {k: sorted(v) for k, v in rfe._final_features.items()}


def test_device_resident_extractor_random_graph_equals_pandas_facing_extractor():
    from graphrole_b200.features.device import DeviceRecursiveFeatureExtractor
    G = nx.gnm_random_graph(2000, 9000, seed=11)
    a = RecursiveFeatureExtractor(G).extract_features()
    csr = interface.get_interface(G)(G).to_csr()
    rfe = DeviceRecursiveFeatureExtractor(csr, device=DEV)
    b = rfe.extract_features()
    assert list(a.columns) == list(b.columns)
    np.testing.assert_allclose(a.values.astype(float), b.values, rtol=1e-6)
    assert set(rfe.timings_ms) >= {'level0', 'aggregate', 'bin', 'pairwise'}
