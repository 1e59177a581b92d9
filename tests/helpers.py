"""Shared test helpers: rebuild graphs / frames from the golden fixtures."""
import networkx as nx
import numpy as np
import pandas as pd


def graph_from_json(g, node_attrs=None):
    G = nx.DiGraph() if g['directed'] else nx.Graph()
    G.add_nodes_from(g['nodes'])
    for u, v, w in g['edges']:
        if g['weighted']:
            G.add_edge(u, v, weight=w)
        else:
            G.add_edge(u, v)
    if node_attrs:
        for node, attrs in node_attrs.items():
            key = node if node in G.nodes else int(node)
            G.nodes[key].update(attrs)
    return G


def frame_from_json(f):
    return pd.DataFrame(np.array(f['values'], dtype=float).reshape(len(f['index']),
                                                                  len(f['columns'])),
                        index=f['index'], columns=f['columns'])


def random_case(z, name):
    """(directed, n, src, dst, X, out_index, out_columns, out) of a refex_random.npz case."""
    get = lambda k: z[f'{name}__{k}']  # noqa: E731
    return (bool(get('directed')), int(get('n')), get('src'), get('dst'), get('X'),
            get('out_index'), [str(c) for c in get('out_columns')], get('out'))


RANDOM_CASES = ['er_undirected', 'er_directed', 'ba_undirected', 'sparse_with_isolates']

# Path-graph known-answer table, restated from the reference's own test
# (tests/test_features/test_extract.py:109-116)
PATH4_EXPECTED = {
    'external_edges(sum)':  {'a': 2.0, 'b': 1.0, 'c': 2.0, 'd': 1.0},
    'degree(sum)':          {'a': 3.0, 'b': 2.0, 'c': 3.0, 'd': 2.0},
    'internal_edges(sum)':  {'a': 3.0, 'b': 2.0, 'c': 3.0, 'd': 2.0},
    'external_edges(mean)': {'a': 1.0, 'b': 1.0, 'c': 1.0, 'd': 1.0},
    'degree(mean)':         {'a': 1.5, 'b': 2.0, 'c': 1.5, 'd': 2.0},
    'internal_edges(mean)': {'a': 1.5, 'b': 2.0, 'c': 1.5, 'd': 2.0},
}
