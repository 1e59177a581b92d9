"""Parity of hot path A on the GPU (through the C-ABI) against the pinned oracle and the
reference's golden vectors.

Tolerance (north_star): feature matrices within 1e-5 relative in fp32.  Stated here as
|gpu - ref| <= 1e-5 * |ref| + 2^-20 * sum_k |x_k|  -- 1e-5 relative to the value, plus 16 fp32
unit round-offs of the row's absolute sum, which is what bounds any fp32 summation when signed
inputs cancel (for non-negative features the second term is < 1e-6 relative).
Integer-valued inputs (the reference's level-0 features) must come back exact.
"""
import numpy as np
import pandas as pd
import pytest
import torch

from graphrole_b200 import RecursiveFeatureExtractor, _native
from graphrole_b200.graph.csr import CSRGraph
from oracle import refex_oracle as oracle
from helpers import (PATH4_EXPECTED, RANDOM_CASES, frame_from_json, graph_from_json,
                     random_case)

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def gpu_aggregate(g, X, **kw):
    h = g.handle('cuda:0')
    out = h.aggregate(torch.as_tensor(X, dtype=torch.float32, device='cuda:0'), **kw)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def assert_parity(got, rp, ci, X32):
    S, M = oracle.aggregate_csr(rp, ci, X32.astype(np.float64))
    Sabs, Mabs = oracle.aggregate_csr(rp, ci, np.abs(X32).astype(np.float64))
    d = X32.shape[1]
    for part, ref, scale in ((got[:, :d], S, Sabs), (got[:, d:], M, Mabs)):
        err = np.abs(part.astype(np.float64) - ref)
        bound = RTOL * np.abs(ref) + scale * 2.0 ** -20 + 1e-30
        assert (err <= bound).all(), f'max rel err {np.max(err / (np.abs(ref) + 1e-30)):.3e}'


def test_native_library_is_loaded():
    lib = _native.load()
    assert b'sm_100a' in lib.gr_version()
    before = _native.launch_count()
    g = CSRGraph.from_edges([0, 1], [1, 2], n=3)
    gpu_aggregate(g, np.ones((3, 4), dtype=np.float32))
    assert _native.launch_count() > before


@pytest.mark.parametrize('name', RANDOM_CASES)
def test_random_graphs_vs_reference_golden(refex_random, name):
    directed, n, src, dst, X, out_index, out_cols, out = random_case(refex_random, name)
    g = CSRGraph.from_edges(src, dst, n=n, directed=directed)
    got = gpu_aggregate(g, X.astype(np.float32))
    rp, ci = g.host_arrays()
    assert_parity(got, rp, ci, X.astype(np.float32))
    # and directly against what the reference printed (float64 inputs -> allow input rounding)
    np.testing.assert_allclose(got[out_index], out, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize('d', [1, 2, 3, 4, 5, 8, 12, 16, 31, 32, 33, 64, 100, 128, 130, 256, 260])
def test_feature_widths(d):
    rng = np.random.RandomState(d)
    n = 500
    src = rng.randint(0, n, 4000)
    dst = rng.randint(0, n, 4000)
    g = CSRGraph.from_edges(src, dst, n=n, directed=True)
    X = rng.rand(n, d).astype(np.float32)
    got = gpu_aggregate(g, X)
    rp, ci = g.host_arrays()
    assert_parity(got, rp, ci, X)


def test_strided_input_and_row_ranges():
    rng = np.random.RandomState(0)
    n, d = 700, 24
    g = CSRGraph.from_edges(rng.randint(0, n, 9000), rng.randint(0, n, 9000), n=n)
    wide = torch.as_tensor(rng.rand(n, 64).astype(np.float32), device='cuda:0')
    rp, ci = g.host_arrays()
    for c0 in (0, 4, 5):       # 5: misaligned -> scalar path
        X = wide[:, c0:c0 + d]
        full = g.handle('cuda:0').aggregate(X).cpu().numpy()
        assert_parity(full, rp, ci, X.cpu().numpy())
        for lo, hi in ((0, 1), (13, 14), (100, 555), (699, 700), (0, 700), (7, 7)):
            part = g.handle('cuda:0').aggregate(X, row_lo=lo, row_hi=hi).cpu().numpy()
            np.testing.assert_array_equal(part, full[lo:hi])


def test_integer_features_are_exact_and_dangling_rows_zero():
    rng = np.random.RandomState(1)
    n = 300
    src, dst = rng.randint(0, 200, 1500), rng.randint(0, 200, 1500)   # nodes >= 200 isolated
    g = CSRGraph.from_edges(src, dst, n=n)
    X = rng.randint(0, 50, size=(n, 3)).astype(np.float32)
    got = gpu_aggregate(g, X)
    rp, ci = g.host_arrays()
    S, M = oracle.aggregate_csr(rp, ci, X)
    np.testing.assert_array_equal(got[:, :3], S.astype(np.float32))
    np.testing.assert_allclose(got[:, 3:], M, rtol=1e-6)
    assert not got[200:].any()
    assert not np.isnan(got).any()


def test_hub_rows_are_split_and_deterministic():
    """Power-law style: a few rows far above the hub threshold (2048 arcs)."""
    rng = np.random.RandomState(2)
    n, d = 30000, 64
    hubs = np.array([0, 7, 29999])
    src = np.concatenate([np.repeat(hubs, [20000, 5000, 2049]), rng.randint(0, n, 60000)])
    dst = np.concatenate([rng.choice(n, 20000, replace=False), rng.choice(n, 5000, replace=False),
                          rng.choice(n, 2049, replace=False), rng.randint(0, n, 60000)])
    g = CSRGraph.from_edges(src, dst, n=n, directed=True)
    info = g.handle('cuda:0').info()
    assert info['n_hub_rows'] == 3 and info['n_hub_segments'] == 20 + 5 + 3
    X = (rng.rand(n, d) * 2 - 0.5).astype(np.float32)
    got = gpu_aggregate(g, X)
    rp, ci = g.host_arrays()
    assert_parity(got, rp, ci, X)
    again = gpu_aggregate(g, X)
    np.testing.assert_array_equal(got, again)        # bitwise reproducible
    # hub rows inside / outside a row range
    part = g.handle('cuda:0').aggregate(torch.as_tensor(X, device='cuda:0'), row_lo=5,
                                        row_hi=n).cpu().numpy()
    np.testing.assert_array_equal(part, got[5:])


def test_hot_row_hints_do_not_change_results():
    """GR_CSR_HOT_HINTS only changes cache policy: bit-identical output, hot rows = the rows
    of highest in-degree."""
    from graphrole_b200.graph.generators import barabasi_albert_csr
    g = barabasi_albert_csr(400_000, 8, seed=2, device='cuda:0')
    X = torch.rand(g.n, 64, device='cuda:0')
    plain = _native.CsrHandle(g.rowptr, g.colidx, hot_hints=False)
    hinted = _native.CsrHandle(g.rowptr, g.colidx, hot_hints=True)
    assert plain.info()['n_hot_rows'] == 0
    n_hot = hinted.info()['n_hot_rows']
    assert 0 < n_hot <= (48 << 20) // 256
    a, b = plain.aggregate(X), hinted.aggregate(X)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    indeg = torch.bincount(g.colidx.long(), minlength=g.n)
    kth = torch.sort(indeg, descending=True).values[n_hot - 1]
    assert int((indeg > kth).sum()) <= n_hot


def test_empty_graph_rows_and_invalid_inputs():
    g = CSRGraph(np.zeros(6, dtype=np.int64), np.zeros(0, dtype=np.int32))
    got = gpu_aggregate(g, np.ones((5, 4), dtype=np.float32))
    assert got.shape == (5, 8) and not got.any()
    bad = CSRGraph(np.array([0, 2, 1]), np.array([0, 1], dtype=np.int32))
    with pytest.raises(ValueError):
        bad.handle('cuda:0')
    bad = CSRGraph(np.array([0, 1, 2]), np.array([0, 5], dtype=np.int32))
    with pytest.raises(ValueError):
        bad.handle('cuda:0')
    good = CSRGraph.from_edges([0], [1], n=2)
    with pytest.raises(ValueError):
        good.handle('cuda:0').aggregate(torch.ones(3, 4, device='cuda:0'))


def test_host_buffer_entry_point_matches_device_path():
    rng = np.random.RandomState(3)
    n, d, levels = 4000, 32, 3
    g = CSRGraph.from_edges(rng.randint(0, n, 40000), rng.randint(0, n, 40000), n=n)
    X = torch.from_numpy(rng.rand(n, d).astype(np.float32)).pin_memory()
    h = g.handle('cuda:0')
    out = h.levels_host(X, levels, recurse_on='mean').numpy()
    cur = X.cuda()
    for l in range(levels):
        lvl = h.aggregate(cur)
        np.testing.assert_array_equal(out[l], lvl.cpu().numpy())
        cur = lvl[:, d:]
    rp, ci = g.host_arrays()
    assert_parity(out[0], rp, ci, X.numpy())


# ---- the reference's API surface on top of the kernel -----------------------------------

def seeded(G, **kw):
    rfe = RecursiveFeatureExtractor(G, aggs=[np.sum, np.mean], **kw)
    rfe._features = rfe.graph.get_neighborhood_features()
    rfe._final_features = {0: rfe._features.to_dict()}
    rfe.generation_count = 1
    return rfe


def test_get_next_features_known_answer(refex_cases):
    """Reference KAT, tests/test_features/test_extract.py:104-122, through the GPU."""
    G = graph_from_json(refex_cases['path4']['graph'])
    got = seeded(G)._get_next_features()
    expected = pd.DataFrame(PATH4_EXPECTED)
    assert np.allclose(got.sort_index(axis=1).sort_index(axis=0).values,
                       expected.sort_index(axis=1).sort_index(axis=0).values)
    ref = frame_from_json(refex_cases['path4']['next'])
    assert list(got.columns) == list(ref.columns) and list(got.index) == list(ref.index)


@pytest.mark.parametrize('name', ['dangling', 'directed_weighted', 'undirected_weighted',
                                  'attributes'])
def test_get_next_features_small_cases(refex_cases, name):
    case = refex_cases[name]
    G = graph_from_json(case['graph'], case.get('node_attrs'))
    kw = {'attributes': True} if name == 'attributes' else {}
    got = seeded(G, **kw)._get_next_features()
    ref = frame_from_json(case['next'])
    assert list(got.columns) == list(ref.columns)
    assert list(got.index) == list(ref.index)
    np.testing.assert_allclose(got.values, ref.values, rtol=RTOL, atol=1e-12)
    assert got.notnull().all().all()


@pytest.mark.parametrize('name', ['dangling', 'directed_weighted', 'undirected_weighted',
                                  'karate', 'karate_weighted'])
def test_extract_features_end_to_end(refex_cases, name):
    case = refex_cases[name]
    G = graph_from_json(case['graph'])
    rfe = RecursiveFeatureExtractor(G)          # default aggs work here (pandas-3 safe)
    got = rfe.extract_features()
    ref = frame_from_json(case['features'])
    assert rfe.generation_count == case['generation_count']
    assert list(got.columns) == list(ref.columns)
    assert list(got.index) == list(ref.index)
    np.testing.assert_allclose(got.values.astype(float), ref.values, rtol=RTOL, atol=1e-12)
    if name == 'karate':
        assert {str(k): sorted(v) for k, v in rfe._final_features.items()} == \
            case['retained_by_generation']
        nb = case['notebook_table']
        np.testing.assert_allclose(got[nb['columns']].values, np.array(nb['values']),
                                   atol=5.1e-7 + 1e-5)
    # memoised second call (tests/test_features/test_extract.py:210-214)
    pd.testing.assert_frame_equal(got, rfe.extract_features())


def test_missing_rows_are_skipped_like_pandas_skipna():
    G = graph_from_json({'directed': False, 'nodes': [0, 1, 2, 3], 'weighted': False,
                         'edges': [[0, 1, 1], [0, 2, 1], [2, 3, 1]]})
    rfe = seeded(G)
    rfe._features = rfe._features.drop(index=[1])    # neighbour without a feature row
    got = rfe._get_next_features()
    feats = rfe._features
    exp = oracle.pandas_chain_rows(feats.reindex(range(4)), feats.columns, range(4),
                                   *rfe.graph.to_csr().host_arrays())
    np.testing.assert_allclose(got.values, exp.values, rtol=RTOL)


@pytest.mark.parametrize('n,deg,d', [(200_000, 16, 32), (100_000, 40, 64)])
def test_larger_random_graph_properties(n, deg, d):
    """Sizes the oracle still finishes in seconds, plus size-independent properties:
    linearity, mean*deg == sum, and column sums (a checksum of checksums)."""
    from graphrole_b200.graph.generators import erdos_renyi_csr
    g = erdos_renyi_csr(n, n * deg // 2, seed=5, device='cuda:0')
    h = g.handle('cuda:0')
    gen = torch.Generator(device='cuda:0').manual_seed(1)
    X = torch.rand(n, d, device='cuda:0', generator=gen)
    Y = torch.rand(n, d, device='cuda:0', generator=gen)
    ox, oy, oxy = h.aggregate(X), h.aggregate(Y), h.aggregate(X + 2 * Y)
    torch.testing.assert_close(oxy, ox + 2 * oy, rtol=1e-5, atol=1e-5)
    degs = g.out_degree().to(torch.float32)[:, None]
    torch.testing.assert_close(ox[:, d:] * degs, ox[:, :d], rtol=1e-5, atol=1e-6)
    # sum over rows of S = sum over nodes of indegree * x (undirected: = outdegree)
    lhs = ox[:, :d].double().sum(0)
    rhs = (X.double() * degs.double()).sum(0)
    torch.testing.assert_close(lhs, rhs, rtol=1e-6, atol=0)
    rp, ci = g.host_arrays()
    rows = np.random.RandomState(0).choice(n, 3000, replace=False)
    S, M = oracle.aggregate_rows_c(rows, rp, ci, X.cpu().numpy())
    got = ox.cpu().numpy()[rows]
    np.testing.assert_allclose(got[:, :d], S, rtol=RTOL)
    np.testing.assert_allclose(got[:, d:], M, rtol=RTOL)


# ---- BASELINE.json's own configurations, at size ----------------------------------------------

def recursion_against_float64(g, d, levels, seed=0):
    """fp32 GPU recursion (schedule alpha: recurse on the mean block) against the float64
    recursion the reference would carry (extract.py:77-83), EVERY row of EVERY level, through
    oracle.level_check_f64.  Returns the per-level (sum, mean) max relative errors."""
    rp, ci = g.host_arrays()
    h = g.handle('cuda:0')
    X = torch.rand(g.n, d, device='cuda:0', generator=torch.Generator('cuda:0').manual_seed(seed))
    X64 = X.cpu().numpy().astype(np.float64)
    deg = g.out_degree().to(torch.float32)[:, None]
    errs = []
    cur = X
    out = None
    for level in range(levels):
        nxt = h.aggregate(cur)
        # the mean is the correctly rounded fp32 quotient of the fp32 sum (never sum * (1 / deg))
        assert torch.equal(nxt[:, d:], torch.where(deg > 0, nxt[:, :d] / deg.clamp(min=1),
                                                    torch.zeros_like(nxt[:, :d])))
        X64, err_sum, err_mean, bad_zero = oracle.level_check_f64(rp, ci, X64, nxt.cpu().numpy())
        assert bad_zero == 0
        errs.append((err_sum, err_mean))
        out, cur = nxt, nxt[:, d:]
        del nxt
    return errs, out


def test_config2_erdos_renyi_1m_20m_four_levels_every_row():
    """BASELINE.json configs[1]: ER |V| = 1 M, |E| = 20 M, 32 features, 4 levels -- whole graph,
    every row, float64 recursion; north_star tolerance 1e-5 relative."""
    from graphrole_b200.graph.generators import erdos_renyi_csr
    g = erdos_renyi_csr(1_000_000, 20_000_000, seed=0, device='cuda:0')
    assert g.n == 1_000_000 and 39_900_000 < g.nnz <= 40_000_000
    errs, _ = recursion_against_float64(g, 32, 4)
    print('C2 max relative error per level (sum, mean):', errs)
    assert max(max(e) for e in errs) <= RTOL


def test_config3_barabasi_albert_10m_200m_five_levels_every_row():
    """BASELINE.json configs[2]: BA |V| = 10 M, |E| ~ 200 M (nnz ~ 400 M, max degree ~ 9e4: hub
    rows, 64-bit arc offsets past 2^31 bytes), 64 features, 5 levels -- whole graph, every row of
    every level against the float64 recursion."""
    import psutil
    from graphrole_b200.graph.generators import barabasi_albert_csr
    free_gpu, _ = torch.cuda.mem_get_info(0)
    if psutil.virtual_memory().available < 40e9 or free_gpu < 40e9:
        pytest.skip('needs ~25 GB of host memory and ~25 GB of HBM')
    g = barabasi_albert_csr(10_000_000, 20, seed=0, device='cuda:0')
    assert g.nnz > 399_000_000
    info = g.handle('cuda:0').info()
    assert info['n_hub_rows'] > 100
    errs, last = recursion_against_float64(g, 64, 5)
    print('C3 max relative error per level (sum, mean):', errs)
    assert max(max(e) for e in errs) <= RTOL
    # checksum of checksums on the last level: sum over rows of S = sum_j indeg(j) x_j
    del last


def test_mean_keeps_exact_ties_like_float64_division():
    """a degree-7 node whose neighbours all carry the same integer must get exactly that
    integer as its mean (21 * fl(1/7) = 3.0000002 would split a tie the reference keeps, and
    vertical_log_binning is rank based).  All (degree, value) pairs up to 200 x 200."""
    degs = np.arange(1, 201)
    rows = np.repeat(np.arange(200), degs)                 # node i has degree i + 1
    cols = 200 + np.concatenate([np.arange(k) for k in degs])
    g = CSRGraph.from_edges(rows, cols, n=400, directed=True)
    vals = np.arange(1, 201, dtype=np.float32)
    X = np.tile(vals, (400, 1))                            # column c holds the value c + 1
    got = gpu_aggregate(g, X)
    np.testing.assert_array_equal(got[:200, 200:], np.tile(vals, (200, 1)))
    np.testing.assert_array_equal(got[:200, :200], degs[:, None].astype(np.float32) * vals[None, :])
    # hub rows follow the same rule
    hub = CSRGraph.from_edges(np.zeros(2051, dtype=np.int64), 1 + np.arange(2051), n=2052,
                              directed=True)
    Xh = np.full((2052, 4), 3.0, dtype=np.float32)
    out = gpu_aggregate(hub, Xh)
    assert (out[0, 4:] == 3.0).all() and (out[0, :4] == 3.0 * 2051).all()


# ---- node-range shards and the fused gather + broadcast kernel -------

def test_row_slice_shards_reproduce_the_unsharded_bits():
    """A shard laid out by CSRGraph.row_slice (colidx slice starting on a 32-arc boundary) must
    give bit-identical rows: sharding may not change results."""
    from graphrole_b200.graph.generators import barabasi_albert_csr
    from graphrole_b200.shard import nnz_balanced_ranges
    g = barabasi_albert_csr(120_000, 9, seed=4, device='cuda:0')     # has hub rows
    X = torch.rand(g.n, 64, device='cuda:0') * 2 - 0.5
    full = g.handle('cuda:0').aggregate(X)
    for world in (2, 3, 8):
        for lo, hi in nnz_balanced_ranges(g.rowptr, world):
            shard = g.row_slice(lo, hi)
            part = shard.handle('cuda:0').aggregate(X)
            assert torch.equal(part, full[lo:hi]), (world, lo, hi)


@pytest.mark.parametrize('d,n_rep', [(64, 1), (64, 3), (64, 8), (32, 2), (12, 5), (7, 3), (128, 4)])
def test_fused_broadcast_kernel_matches_plain_kernel(d, n_rep):
    """gr_refex_aggregate_bcast_f32 with every replica on this GPU: each replica receives the
    shard's mean rows at the shard's global row offset, nothing else is touched, and the values
    are bit-identical to gr_refex_aggregate_f32 (hub rows included)."""
    from graphrole_b200.graph.generators import barabasi_albert_csr
    g = barabasi_albert_csr(300_000, 10, seed=6, device='cuda:0')   # first rows are hubs
    X = torch.rand(g.n, d, device='cuda:0')
    full = g.handle('cuda:0').aggregate(X)
    lo, hi = 1, 141_234
    shard = g.row_slice(lo, hi)
    assert shard.handle('cuda:0').info()['n_hub_rows'] > 0
    reps = [torch.full((g.n, d), -7.0, device='cuda:0') for _ in range(n_rep)]
    sums = torch.empty((hi - lo, d), device='cuda:0')
    shard.handle('cuda:0').aggregate_bcast(X, sums, [r.data_ptr() for r in reps], d, lo)
    torch.cuda.synchronize()
    assert torch.equal(sums, full[lo:hi, :d])
    for r in reps:
        assert torch.equal(r[lo:hi], full[lo:hi, d:])
        assert bool((r[:lo] == -7.0).all()) and bool((r[hi:] == -7.0).all())


def test_fused_broadcast_rejects_bad_arguments():
    g = CSRGraph.from_edges([0, 1], [1, 2], n=3)
    h = g.handle('cuda:0')
    X = torch.ones(3, 4, device='cuda:0')
    with pytest.raises(ValueError):
        h.aggregate_bcast(X, None, [], 4, 0)                       # no replicas
    with pytest.raises(ValueError):
        h.aggregate_bcast(X, None, [X.data_ptr()], 4, 0)           # replica aliases X
    with pytest.raises(ValueError):
        h.aggregate_bcast(X, None, [X.data_ptr() + 4096] * 17, 4, 0)   # more than 16 replicas


def test_peer_barrier_single_rank_and_timeout_flag():
    """The flag barrier with one rank completes immediately; waiting for an epoch nobody
    publishes gives up after the timeout and reports it instead of hanging."""
    words = _native.peer_flag_words()
    buf = _native.PeerBuffer(8 * words, 0)
    flags = buf.tensor((words,), torch.int64)
    flags.zero_()
    torch.cuda.synchronize()
    _native.peer_barrier([buf.ptr], 0, 1)
    torch.cuda.synchronize()
    assert int(flags[0]) == 1 and _native.peer_barrier_timed_out(buf.ptr) == 0
    # two "ranks" sharing one flag array: slot 1 never gets epoch 2 -> timeout path
    _native.peer_barrier([buf.ptr, buf.ptr], 0, 2, timeout_s=0.05)
    torch.cuda.synchronize()
    assert _native.peer_barrier_timed_out(buf.ptr) == 2
    del flags
    buf.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs on one box')
def test_two_gpu_sharded_recursion_is_bit_identical():
    """torchrun, one rank per GPU: fused peer-store exchange and the all-gather exchange both
    reproduce the single-GPU recursion bit for bit (tools/check_sharded.py)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run(
        [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
         '--master-addr', '127.0.0.1', '--master-port', '29731',
         os.path.join(root, 'tools', 'check_sharded.py'), '--size', '200000', '--depth', '3'],
        capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and 'SHARDED CHECK OK' in res.stdout, res.stdout + res.stderr
