#!/usr/bin/env python
"""Generate the committed golden fixtures by running the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    PYTHONPATH=/root/reference python tests/golden/make_golden.py

The reference is imported, never copied.  pandas here is 3.x while the reference pins
pandas<2, so its default ``aggs=[pd.DataFrame.sum, pd.DataFrame.mean]`` raise TypeError;
every call below passes ``aggs=['sum', 'mean']`` (same output names, see
/root/reference/graphrole/features/extract.py:158-162).

Outputs (all small, committed):
  refex_cases.json     path-graph KAT, dangling nodes, directed / self-loop cases, karate
                       end-to-end incl. per-generation retained sets and the notebook table
  refex_random.npz     seeded random graphs with injected float feature matrices and the
                       reference's `_get_next_features` output for them
  prune_cases.json     reference pruner outputs (binning vectors, dropped sets) on seeded data
  prune_reference_tables.json  the reference's own known-answer tables for the pruner, re-run
  rolx_cases.npz       the reference's `encode` (= sklearn KMeans(random_state=1) on the flattened
                       matrix) and description-length costs on seeded inputs, plus one row of
                       the model-selection grid (fixed factors, bits 1..8)
  nmf_cases.npz        sklearn MU runs with explicit (W0, H0): factors after a fixed number
                       of iterations and at the stopping iteration, plus
                       graphrole.roles.factor.get_nmf_decomposition under np.random.seed
"""
import io
import json
import os
import sys
import warnings

import networkx as nx
import numpy as np
import pandas as pd

REF = os.environ.get('GRAPHROLE_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, 'examples'))

from graphrole import RecursiveFeatureExtractor, RoleExtractor  # noqa: E402
from graphrole.features.prune import FeaturePruner, vertical_log_binning  # noqa: E402
from graphrole.roles import factor as ref_factor  # noqa: E402
from graphrole.roles.description_length import get_description_length_costs  # noqa: E402
from data.data import load_nx_karate_club_graph  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
AGGS = ['sum', 'mean']


def frame_to_json(df):
    return {
        'index': [x if isinstance(x, str) else int(x) for x in df.index],
        'columns': list(df.columns),
        'values': df.values.astype(float).tolist(),
    }


def seeded_next_features(G, **kwargs):
    """The reference tests' own injection seam (tests/test_features/test_extract.py:87-92)."""
    rfe = RecursiveFeatureExtractor(G, aggs=AGGS, **kwargs)
    rfe._features = rfe.graph.get_neighborhood_features()
    rfe._final_features = {0: rfe._features.to_dict()}
    rfe.generation_count = 1
    return rfe._features, rfe._get_next_features()


def graph_to_json(G):
    return {
        'directed': G.is_directed(),
        'nodes': [x if isinstance(x, str) else int(x) for x in G.nodes],
        'edges': [[u if isinstance(u, str) else int(u), v if isinstance(v, str) else int(v),
                   float(d.get('weight', 1))] for u, v, d in G.edges(data=True)],
        'weighted': any('weight' in d for _, _, d in G.edges(data=True)),
    }


def refex_cases():
    cases = {}

    # (1) path graph, tests/test_features/test_extract.py:79-80,104-122
    G = nx.Graph([('a', 'b'), ('a', 'c'), ('c', 'd')])
    lvl0, nxt = seeded_next_features(G)
    cases['path4'] = {'graph': graph_to_json(G), 'level0': frame_to_json(lvl0),
                      'next': frame_to_json(nxt)}

    # (2) dangling nodes, tests/test_features/test_extract.py:36-67
    G = nx.Graph()
    G.add_nodes_from(['a', 'b', 'c', 'd'])
    G.add_edge('a', 'c')
    lvl0, nxt = seeded_next_features(G)
    rfe = RecursiveFeatureExtractor(G, aggs=AGGS)
    full = rfe.extract_features()
    cases['dangling'] = {'graph': graph_to_json(G), 'level0': frame_to_json(lvl0),
                         'next': frame_to_json(nxt), 'features': frame_to_json(full),
                         'generation_count': rfe.generation_count}

    # (3) directed, weighted, with a self loop and unsorted integer labels
    G = nx.DiGraph()
    G.add_weighted_edges_from([(5, 2, 2.0), (2, 9, 1.0), (9, 5, 3.0), (1, 5, 1.0), (5, 5, 4.0),
                               (9, 1, 2.5), (7, 1, 1.0)])
    G.add_node(12)
    lvl0, nxt = seeded_next_features(G)
    rfe = RecursiveFeatureExtractor(G, aggs=AGGS)
    full = rfe.extract_features()
    cases['directed_weighted'] = {'graph': graph_to_json(G), 'level0': frame_to_json(lvl0),
                                  'next': frame_to_json(nxt), 'features': frame_to_json(full),
                                  'generation_count': rfe.generation_count}

    # (4) undirected weighted graph from the interface tests' shape (7 nodes), with a self loop
    G = nx.Graph()
    G.add_weighted_edges_from([(0, 1, 2.0), (0, 2, 1.0), (1, 2, 0.5), (2, 3, 4.0), (3, 4, 1.0),
                               (4, 5, 1.5), (5, 6, 1.0), (6, 4, 2.0), (3, 3, 7.0)])
    lvl0, nxt = seeded_next_features(G)
    rfe = RecursiveFeatureExtractor(G, aggs=AGGS)
    full = rfe.extract_features()
    cases['undirected_weighted'] = {'graph': graph_to_json(G), 'level0': frame_to_json(lvl0),
                                    'next': frame_to_json(nxt), 'features': frame_to_json(full),
                                    'generation_count': rfe.generation_count}

    # (5) karate club end to end, examples/example.ipynb cell 3 (+ per-generation bookkeeping)
    G = load_nx_karate_club_graph(weighted=False)
    rfe = RecursiveFeatureExtractor(G, aggs=AGGS)
    full = rfe.extract_features()
    gens = {str(g): sorted(d.keys()) for g, d in rfe._final_features.items()}
    lvl0 = RecursiveFeatureExtractor(G, aggs=AGGS).graph.get_neighborhood_features()
    cases['karate'] = {'graph': graph_to_json(G), 'level0': frame_to_json(lvl0),
                       'features': frame_to_json(full), 'generation_count': rfe.generation_count,
                       'retained_by_generation': gens,
                       'notebook_table': notebook_feature_table()}

    # (6) karate with weights (level-0 weighted degree semantics feeding the recursion)
    G = load_nx_karate_club_graph(weighted=True)
    rfe = RecursiveFeatureExtractor(G, aggs=AGGS)
    full = rfe.extract_features()
    lvl0 = RecursiveFeatureExtractor(G, aggs=AGGS).graph.get_neighborhood_features()
    cases['karate_weighted'] = {'graph': graph_to_json(G), 'level0': frame_to_json(lvl0),
                                'features': frame_to_json(full),
                                'generation_count': rfe.generation_count}

    # (7) node attributes as base features (README.md:69-92)
    G = nx.Graph([(0, 1), (1, 2), (2, 0), (2, 3), (3, 4)])
    rng = np.random.RandomState(3)
    for node in G.nodes:
        G.nodes[node]['score'] = float(rng.rand())
        G.nodes[node]['label'] = 'x'      # non-numeric: ignored
    G.nodes[4]['extra'] = 2.0
    lvl0, nxt = seeded_next_features(G, attributes=True)
    cases['attributes'] = {'graph': graph_to_json(G),
                           'node_attrs': {str(n): {k: v for k, v in d.items()}
                                          for n, d in G.nodes(data=True)},
                           'level0': frame_to_json(lvl0), 'next': frame_to_json(nxt)}

    # (8) the graphs of the reference's own interface tests (tests/test_graph/
    # test_interface.py:50-67: 7 nodes, 7 edges, weights, node attributes); the level-0 tables
    # the reference asserts there (:124-148, :150-186, :188-220) are what these frames hold
    edges = [(0, 1), (0, 2), (0, 3), (3, 6), (4, 5), (4, 6), (5, 6)]
    weights = [2, 1.5, 3, 0.25, 0.75, 2.5, 1]
    G = nx.Graph()
    G.add_nodes_from(range(7))
    G.add_edges_from(edges)
    lvl0 = RecursiveFeatureExtractor(G, aggs=AGGS).graph.get_neighborhood_features()
    assert lvl0['internal_edges'].tolist() == [3, 1, 1, 2, 3, 3, 4]          # :133-138
    assert lvl0['external_edges'].tolist() == [1, 2, 2, 4, 1, 1, 1]          # :139-144
    cases['iface_undirected'] = {'graph': graph_to_json(G), 'level0': frame_to_json(lvl0)}
    G = nx.DiGraph()
    G.add_nodes_from(range(7))
    G.add_weighted_edges_from([(u, v, w) for (u, v), w in zip(edges, weights)])
    lvl0 = RecursiveFeatureExtractor(G, aggs=AGGS).graph.get_neighborhood_features()
    assert lvl0['internal_edges'].tolist() == [6.5, 0, 0, 0.25, 4.25, 1, 0]  # :172-177
    assert lvl0['external_edges'].tolist() == [0.25, 0, 0, 0, 0, 0, 0]       # :178-183
    cases['iface_directed_weighted'] = {'graph': graph_to_json(G), 'level0': frame_to_json(lvl0)}

    with open(os.path.join(HERE, 'refex_cases.json'), 'w') as f:
        json.dump(cases, f, indent=1)


def notebook_feature_table():
    """The printed 34x7 table stored in examples/example.ipynb (cell 3 stdout), parsed as data."""
    nb = json.load(open(os.path.join(REF, 'examples', 'example.ipynb')))
    text = None
    for cell in nb['cells']:
        if cell['cell_type'] != 'code':
            continue
        for out in cell.get('outputs', []):
            t = ''.join(out.get('text', []))
            if 'Features extracted from' in t:
                text = t
    assert text is not None
    head, body = text.split('\n', 2)[1:] if text.startswith('\n') else text.split('\n', 1)
    n_gen = int(head.split('from')[1].split('recursive')[0])
    # pandas wraps wide frames into blocks separated by blank lines; a trailing '\' marks a wrap
    blocks = [b for b in body.split('\n\n') if b.strip()]
    frames = []
    for b in blocks:
        lines = [ln.rstrip('\\').rstrip() for ln in b.split('\n') if ln.strip()]
        frames.append(pd.read_csv(io.StringIO('\n'.join(lines)), sep=r'\s{2,}', engine='python',
                                  index_col=0))
    table = pd.concat(frames, axis=1)
    return {'generations': n_gen, **frame_to_json(table)}


def refex_random():
    """Seeded random graphs + injected float features through the sanctioned seam."""
    out = {}
    specs = [
        ('er_undirected', nx.gnm_random_graph(300, 3000, seed=1), 5),
        ('er_directed', nx.gnm_random_graph(200, 1500, seed=2, directed=True), 4),
        ('ba_undirected', nx.barabasi_albert_graph(400, 3, seed=3), 7),
        ('sparse_with_isolates', nx.gnm_random_graph(250, 120, seed=4), 3),
    ]
    for name, G, d in specs:
        if name == 'er_undirected':
            G.add_edge(0, 0)
            G.add_edge(17, 17)
        rng = np.random.RandomState(abs(hash(name)) % (2 ** 31) if False else len(name))
        nodes = sorted(G.nodes)
        X = rng.rand(len(nodes), d) * 10.0 - (2.0 if 'directed' in name else 0.0)
        cols = [f'f{j}' for j in range(d)]
        rfe = RecursiveFeatureExtractor(G, aggs=AGGS)
        rfe._features = pd.DataFrame(X, index=nodes, columns=cols)
        rfe._final_features = {0: rfe._features.to_dict()}
        rfe.generation_count = 1
        nxt = rfe._get_next_features()
        # dense row layout in sorted-label order; nodes without out-neighbours are absent
        # from the reference output under pandas>=2 only if reindex yields empty -> they are
        # present here with zeros after fillna; record index explicitly.
        src, dst = zip(*[(u, v) for u, v in G.edges]) if G.number_of_edges() else ((), ())
        out[f'{name}__directed'] = np.array(G.is_directed())
        out[f'{name}__n'] = np.array(len(nodes))
        out[f'{name}__src'] = np.array(src, dtype=np.int64)
        out[f'{name}__dst'] = np.array(dst, dtype=np.int64)
        out[f'{name}__X'] = X
        out[f'{name}__out_index'] = np.array(list(nxt.index), dtype=np.int64)
        out[f'{name}__out_columns'] = np.array(list(nxt.columns))
        out[f'{name}__out'] = nxt.values.astype(np.float64)
    np.savez_compressed(os.path.join(HERE, 'refex_random.npz'), **out)


def prune_cases():
    cases = {'binning': [], 'prune': []}
    rng = np.random.RandomState(11)
    vecs = [
        rng.randint(0, 5, size=20).astype(float),
        rng.rand(33),
        np.zeros(7),
        np.arange(16, dtype=float),
        np.concatenate([np.zeros(10), rng.rand(6)]),
        rng.randint(0, 3, size=50).astype(float),
    ]
    for v in vecs:
        for frac in (0.5, 0.25, 0.8):
            cases['binning'].append({'arr': v.tolist(), 'frac': frac,
                                     'binned': vertical_log_binning(v, frac).tolist()})
    for seed in range(6):
        rng = np.random.RandomState(100 + seed)
        n = 40
        base = rng.randint(0, 6, size=(n, 3)).astype(float)
        feats = pd.DataFrame({
            'a': base[:, 0], 'b': base[:, 1], 'c': base[:, 2],
            'a2': base[:, 0] * 2.0,                     # same ranks as a -> same bins
            'b_noise': base[:, 1] + (rng.rand(n) < 0.1),
            'd': rng.rand(n),
        })
        gen_dict = {0: {'a': {}, 'b': {}, 'c': {}}, 1: {'a2': {}, 'b_noise': {}, 'd': {}}}
        for thresh in (0, 1, 2):
            dropped = FeaturePruner(gen_dict, thresh).prune_features(feats)
            cases['prune'].append({'values': feats.values.tolist(), 'columns': list(feats.columns),
                                   'generations': {str(k): sorted(v) for k, v in gen_dict.items()},
                                   'thresh': thresh, 'dropped': sorted(dropped)})
    with open(os.path.join(HERE, 'prune_cases.json'), 'w') as f:
        json.dump(cases, f)


def prune_reference_tables():
    """The inputs of the reference's own known-answer tables for the pruning step
    (tests/test_features/test_prune.py:17-99 vertical_log_binning, :119-153 prune_features,
    :155-185 _group_features), re-run through the unmodified reference: pins our restatements
    to exactly the cases the reference pins itself on."""
    r10 = np.arange(10)
    binning_inputs = {
        'empty': ([], 0.5), 'single 0': ([0], 0.5), 'single nonzero': ([1], 0.5),
        'repeated': ([1, 1], 0.5), '2 bins': ([1, 2], 0.5),
        '2 bins with repeated lower bin': ([1, 2, 1], 0.5),
        '2 bins with repeated upper bin': ([1, 2, 2], 0.5),
        'negative and zeros': ([-1, 0, 0], 0.5), '1 through 4': ([1, 2, 3, 4], 0.5),
        '1 through 5': ([1, 2, 3, 4, 5], 0.5), '1 through 6': ([1, 2, 3, 4, 5, 6], 0.5),
        'range(10)': (r10.tolist(), 0.5), '-range(10)': ((-1 * r10).tolist(), 0.5),
        'non-integer': ((-0.1 * r10).tolist(), 0.5), 'frac=0.1': (r10.tolist(), 0.1),
        'frac=0.25': (r10.tolist(), 0.25),
    }
    out = {'binning': [], 'prune': [], 'group': []}
    for name, (arr, frac) in binning_inputs.items():
        binned = vertical_log_binning(np.array(arr), frac=frac)
        out['binning'].append({'name': name, 'arr': arr, 'frac': frac,
                               'binned': np.asarray(binned).tolist()})
    feats = pd.DataFrame({'a': [1, 2, 3, 10], 'b': [1, 2, 3, 1], 'c': [2, 1, 1, 4],
                          'd': [1, 1, 1, 1], 'e': [1, 1, 2, 0]})
    gen_dict = {0: {'a': {}, 'b': {}, 'c': {}}, 1: {'d': {}, 'e': {}}}
    for thresh in (0, 1, 2):
        dropped = FeaturePruner(gen_dict, thresh).prune_features(feats)
        out['prune'].append({'values': feats.values.tolist(), 'columns': list(feats.columns),
                             'generations': {str(k): sorted(v) for k, v in gen_dict.items()},
                             'thresh': thresh, 'dropped': sorted(dropped)})
    feats = pd.DataFrame({'a': [1, 2, 3], 'b': [1, 2, 3], 'c': [2, 1, 1], 'd': [1, 1, 1]})
    gen_dict = {0: {'b': {}, 'a': {}}, 1: {'c': {}, 'd': {}}}
    for thresh in (0, 1, 2, -1):
        groups = FeaturePruner(gen_dict, thresh)._group_features(feats)
        out['group'].append({'values': feats.values.tolist(), 'columns': list(feats.columns),
                             'thresh': thresh,
                             'groups': sorted(sorted(g) for g in groups)})
    with open(os.path.join(HERE, 'prune_reference_tables.json'), 'w') as f:
        json.dump(out, f, indent=1)


def nmf_cases():
    """sklearn (the third-party owner of path B's arithmetic; installed 1.9.0) with shared init.

    Reference call site: graphrole/roles/factor.py:19-25 -> NMF(solver='mu', init='nndsvda').
    """
    import sklearn
    from sklearn.decomposition import _nmf as sk

    out = {'sklearn_version': np.array(sklearn.__version__)}
    specs = [('rand20x30', 20, 30, [2, 4, 7]), ('rand300x64', 300, 64, [4, 8, 16]),
             ('planted500x48', 500, 48, [4, 8])]
    for name, n, f, ranks in specs:
        rng = np.random.RandomState(0)
        if name.startswith('planted'):
            X = rng.rand(n, 6) @ rng.rand(6, f) + 0.01 * rng.rand(n, f)
        else:
            X = rng.rand(n, f)
        out[f'{name}__X'] = X
        for r in ranks:
            np.random.seed(1234)
            W0, H0 = sk._initialize_nmf(X, r, init='nndsvda')
            out[f'{name}__r{r}__W0'] = W0
            out[f'{name}__r{r}__H0'] = H0
            # fixed 50 iterations, no convergence test
            W, H, it = sk._fit_multiplicative_update(X, W0.copy(), H0.copy(), 'frobenius',
                                                     max_iter=50, tol=0)
            out[f'{name}__r{r}__W50'] = W
            out[f'{name}__r{r}__H50'] = H
            # defaults of NMF (tol=1e-4, max_iter=200): stopping iteration and factors
            W, H, it = sk._fit_multiplicative_update(X, W0.copy(), H0.copy(), 'frobenius',
                                                     max_iter=200, tol=1e-4)
            out[f'{name}__r{r}__Wconv'] = W
            out[f'{name}__r{r}__Hconv'] = H
            out[f'{name}__r{r}__n_iter'] = np.array(it)
            out[f'{name}__r{r}__err'] = np.array(
                sk._beta_divergence(X, W, H, 2, square_root=True))
            # the reference wrapper itself under a seeded global RNG
            np.random.seed(1234)
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                G_ref, F_ref = ref_factor.get_nmf_decomposition(X, r)
            out[f'{name}__r{r}__Gref'] = G_ref
            out[f'{name}__r{r}__Fref'] = F_ref
    np.savez_compressed(os.path.join(HERE, 'nmf_cases.npz'), **out)


def roles_cases():
    """RoleExtractor on the reference's own seeded test input (tests/test_roles/test_extract.py)."""
    np.random.seed(0)
    X = np.random.rand(20, 30)
    feats = pd.DataFrame(X)
    cases = {'X': X.tolist()}
    rows = []
    for roles, bits in [(2, 2), (3, 4), (5, 3)]:
        np.random.seed(7)
        model = RoleExtractor._get_encoded_role_factors(feats, roles, bits)
        enc, err = get_description_length_costs(feats, model)
        rows.append({'roles': roles, 'bits': bits, 'encoding_cost': float(enc),
                     'error_cost': float(err), 'G': model[0].tolist(), 'F': model[1].tolist()})
    cases['encoded'] = rows
    with open(os.path.join(HERE, 'roles_cases.json'), 'w') as f:
        json.dump(cases, f)


def rolx_cases():
    """RolX epilogue: graphrole/roles/factor.py:29-49 (`encode`) and
    graphrole/roles/description_length.py on seeded inputs.  Every encode input holds at least
    as many distinct values as bins (below that scikit-learn's result depends on
    np.argpartition's order of equal keys and is not a stable target).  Grid-valued data with
    almost as many bins as distinct values (round2000x3 at 64 bins: 101 distinct values) are left
    out for the same reason: points exactly half way between two centres are assigned by the
    rounding noise of scikit-learn's BLAS calls."""
    from graphrole.roles import description_length as ref_dl
    out = {}
    rng = np.random.RandomState(0)
    nmf = np.load(os.path.join(HERE, 'nmf_cases.npz'))
    inputs = {
        'rand20x30': (np.random.RandomState(0).rand(20, 30), [2, 4, 8, 16, 64, 256]),
        'exp400x5': (rng.exponential(size=(400, 5)) ** 2, [2, 3, 4, 8, 32, 128, 256]),
        'nmfG300x8': (nmf['rand300x64__r8__Wconv'], [2, 4, 16, 64, 256]),
        'nmfF8x64': (nmf['rand300x64__r8__Hconv'], [2, 4, 16, 64, 256]),
        'round2000x3': (np.round(rng.rand(2000, 3), 2), [2, 4, 8, 16, 32]),
        'halfzeros1000x4': (np.maximum(rng.randn(1000, 4), 0), [2, 4, 8, 16, 64]),
        'vector37': (rng.rand(37), [2, 5, 37]),
    }
    names = []
    for name, (X, bins) in inputs.items():
        out[f'{name}__X'] = X
        out[f'{name}__bins'] = np.array(bins)
        names.append(name)
        for k in bins:
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                out[f'{name}__enc{k}'] = ref_factor.encode(X.copy(), k)
    out['encode_names'] = np.array(names)

    # description length: (V, encoded factors) -> (encoding cost, error cost)
    V = rng.rand(40, 12)
    V[rng.rand(40, 12) < 0.1] = 0.0            # the cost skips v == 0 (description_length.py:57)
    out['dl__V'] = V
    dl_rows = []
    for r, bits in [(2, 1), (3, 3), (5, 5)]:
        np.random.seed(11)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            G, F = ref_factor.get_nmf_decomposition(V, r)
            Ge, Fe = ref_factor.encode(G, 2 ** bits), ref_factor.encode(F, 2 ** bits)
        enc, err = get_description_length_costs(V, (Ge, Fe))
        out[f'dl__r{r}b{bits}__G'] = Ge
        out[f'dl__r{r}b{bits}__F'] = Fe
        out[f'dl__r{r}b{bits}__costs'] = np.array([enc, err])
        assert enc == ref_dl.get_encoding_cost((Ge, Fe))
        dl_rows.append(f'r{r}b{bits}')
    out['dl_names'] = np.array(dl_rows)

    # one row of the model-selection grid (roles/extract.py:121-133) with the factors held fixed:
    # encode + costs for bits 1..8 (NaN where encode raises ValueError)
    V = rng.rand(60, 20) @ np.diag(rng.rand(20) + 0.2)
    np.random.seed(5)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        G, F = ref_factor.get_nmf_decomposition(V, 4)
    grid = np.full((9, 2), np.nan)
    for bits in range(1, 9):
        try:
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                model = (ref_factor.encode(G, 2 ** bits), ref_factor.encode(F, 2 ** bits))
        except ValueError:
            continue
        grid[bits] = get_description_length_costs(V, model)
    out['grid__V'], out['grid__G'], out['grid__F'], out['grid__costs'] = V, G, F, grid

    # the whole model selection on the reference's own test input (tests/test_roles/
    # test_extract.py:81-88, seeded 20 x 30 uniform data).  That test expects 2 roles; under the
    # scikit-learn installed here (1.9.0) the UNMODIFIED reference selects what is recorded below
    # (the reference's examples/example.py:24-28 warns that results depend on the version).
    np.random.seed(0)
    feats = pd.DataFrame(np.random.rand(20, 30))
    rx = RoleExtractor()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        rx.extract_role_factors(feats)
    out['select__X'] = feats.values
    out['select__n_roles'] = np.array(rx.node_role_factor.shape[1])
    out['select__n_levels'] = np.array(len(np.unique(rx.node_role_factor.values)))
    enc_grid, err_grid = np.full((9, 9), np.nan), np.full((9, 9), np.nan)
    for roles in range(2, 9):
        np.random.seed(roles)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            G, F = ref_factor.get_nmf_decomposition(feats.values, roles)
        for bits in range(1, 9):
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter('ignore')
                    model = (ref_factor.encode(G, 2 ** bits), ref_factor.encode(F, 2 ** bits))
            except ValueError:
                continue
            enc_grid[roles, bits], err_grid[roles, bits] = get_description_length_costs(
                feats.values, model)
    out['select__encoding_costs'], out['select__error_costs'] = enc_grid, err_grid
    np.savez_compressed(os.path.join(HERE, 'rolx_cases.npz'), **out)


if __name__ == '__main__':
    if sys.argv[1:] == ['rolx']:        # add this fixture without touching the others
        rolx_cases()
        sys.exit(0)
    refex_cases()
    refex_random()
    prune_cases()
    prune_reference_tables()
    nmf_cases()
    roles_cases()
    rolx_cases()
    print('golden fixtures written to', HERE)
