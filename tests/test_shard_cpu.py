"""Host-side logic of the node-range sharded recursion, exercised with world_size-2 gloo
processes on CPU.  The per-shard aggregation is done by the oracle here (tests may use it);
what is under test is the range computation, the shard slicing and the in-place exchange."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graphrole_b200.graph.generators import barabasi_albert_csr, erdos_renyi_csr
from graphrole_b200.shard import (column_groups, cost_balanced_ranges, default_column_groups,
                                  exchange_rows, nnz_balanced_ranges, time_balanced_ranges)
from oracle import refex_oracle as oracle
from oracle import nmf_oracle


def test_ranges_cover_all_rows_and_balance_arcs():
    g = barabasi_albert_csr(50_000, 10, seed=1, device='cpu')
    for world in (1, 2, 3, 4, 8):
        ranges = nnz_balanced_ranges(g.rowptr, world)
        assert ranges[0][0] == 0 and ranges[-1][1] == g.n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        arcs = [int(g.rowptr[hi] - g.rowptr[lo]) for lo, hi in ranges]
        assert sum(arcs) == g.nnz
        assert max(arcs) - min(arcs) <= 2 * int(g.out_degree().max())


def test_cost_balanced_ranges_trade_arcs_for_rows():
    """Fused exchange: a rank's time is max(arcs, row_cost * rows).  The ranges must cover all
    rows, and their worst rank must beat the arc-balanced split under that cost."""
    g = barabasi_albert_csr(200_000, 10, seed=1, device='cpu')
    rp = g.rowptr

    def worst(ranges, row_cost):
        return max(max(int(rp[hi] - rp[lo]), row_cost * (hi - lo)) for lo, hi in ranges)

    for world in (2, 4, 8):
        row_cost = 10.0 * (world - 1)
        ranges = cost_balanced_ranges(rp, world, row_cost)
        assert len(ranges) == world and ranges[0][0] == 0 and ranges[-1][1] == g.n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        assert all(hi >= lo for lo, hi in ranges)
        assert worst(ranges, row_cost) <= worst(nnz_balanced_ranges(rp, world), row_cost)
        # within one row's worth of the continuous optimum's lower bound
        lower = max(g.nnz / world, row_cost * g.n / world)
        assert worst(ranges, row_cost) <= 1.35 * lower + int(g.out_degree().max())
    # no row cost: the arc-balanced split (up to one row per boundary)
    a = cost_balanced_ranges(rp, 4, 0.0)
    b = nnz_balanced_ranges(rp, 4)
    dmax = int(g.out_degree().max())
    for (alo, ahi), (blo, bhi) in zip(a, b):
        assert abs(int(rp[ahi] - rp[alo]) - int(rp[bhi] - rp[blo])) <= 2 * dmax
    # degenerate inputs
    assert cost_balanced_ranges(torch.zeros(5, dtype=torch.int64), 3, 20.0)[-1][1] == 4
    assert cost_balanced_ranges(rp, 1, 0.0) == [(0, g.n)]


def test_ranges_degenerate():
    rowptr = torch.zeros(5, dtype=torch.int64)
    assert nnz_balanced_ranges(rowptr, 3)[-1][1] == 4
    rowptr = torch.tensor([0, 10, 10, 10])
    ranges = nnz_balanced_ranges(rowptr, 4)
    assert sum(hi - lo for lo, hi in ranges) == 3


def test_row_slice_keeps_global_columns():
    g = erdos_renyi_csr(2000, 10000, seed=2, device='cpu')
    s = g.row_slice(500, 1200)
    rp, ci = g.host_arrays()
    pad = int(rp[500]) % 32      # the slice starts at the 32-arc boundary below the first arc
    assert s.n == 700 and s.n_cols == 2000 and int(s.rowptr[0]) == pad
    assert s.nnz == rp[1200] - rp[500]
    srp, sci = s.host_arrays()
    np.testing.assert_array_equal(sci[pad:], ci[rp[500]:rp[1200]])
    np.testing.assert_array_equal(np.diff(srp), np.diff(rp[500:1201]))
    assert int(srp[-1]) == sci.size


def test_time_balanced_ranges_equalise_a_known_cost():
    """Measured-feedback balancing: with per-range times generated from a hidden cost model
    (arcs + 30 per row, one rank additionally 1.8x slow), two corrections bring max / mean of
    the modelled times from > 1.5 to < 1.05; ranges always tile [0, n)."""
    g = barabasi_albert_csr(200_000, 10, seed=1, device='cpu')
    rp = g.rowptr.double()

    def modelled(ranges):
        return [float(rp[hi] - rp[lo]) + 30.0 * (hi - lo) for lo, hi in ranges]

    for world in (2, 4, 8):
        ranges = nnz_balanced_ranges(g.rowptr, world)
        t = modelled(ranges)
        assert max(t) / (sum(t) / world) > 1.3
        for _ in range(3):
            ranges = time_balanced_ranges(g.rowptr, ranges, modelled(ranges))
            assert ranges[0][0] == 0 and ranges[-1][1] == g.n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        t = modelled(ranges)
        assert max(t) / (sum(t) / world) < 1.05
    # equal times: nothing moves; one range: unchanged
    ranges = nnz_balanced_ranges(g.rowptr, 4)
    same = time_balanced_ranges(g.rowptr, ranges, [1.0] * 4, row_cost=0.0)
    arcs = lambda rs: [int(g.rowptr[hi] - g.rowptr[lo]) for lo, hi in rs]   # noqa: E731
    assert max(abs(a - b) for a, b in zip(arcs(same), arcs(ranges))) <= 2 * int(g.out_degree().max())
    assert time_balanced_ranges(g.rowptr, [(0, g.n)], [3.0]) == [(0, g.n)]


def test_column_groups():
    assert column_groups(64, 1) == [(0, 64)]
    assert column_groups(64, 2) == [(0, 32), (32, 64)]
    assert column_groups(64, 8) == [(i * 8, i * 8 + 8) for i in range(8)]
    assert column_groups(12, 2) == [(0, 4), (4, 12)] or column_groups(12, 2) == [(0, 8), (8, 12)] \
        or column_groups(12, 2) == [(0, 4), (4, 12)]
    for d, c in ((12, 2), (7, 3), (5, 5), (100, 8)):
        groups = column_groups(d, c)
        assert groups[0][0] == 0 and groups[-1][1] == d and len(groups) == c
        assert all(a[1] == b[0] and a[1] > a[0] for a, b in zip(groups, groups[1:]))
    with pytest.raises(ValueError):
        column_groups(3, 4)
    assert default_column_groups(1, 64) == 1 and default_column_groups(4, 64) == 1
    assert default_column_groups(8, 64) == 2
    assert default_column_groups(8, 32) == 1       # 16 columns per group gather too narrowly


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, levels, equal_rows, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = barabasi_albert_csr(3000, 6, seed=3, device='cpu')
        d = 5
        X0 = torch.rand(g.n, d, generator=torch.Generator().manual_seed(0), dtype=torch.float64)
        if equal_rows == 'cost':       # the fused exchange's split: max(arcs, row cost) balanced
            ranges = cost_balanced_ranges(g.rowptr, world, 10.0 * (world - 1))
        elif equal_rows:
            ranges = [(k * g.n // world, (k + 1) * g.n // world) for k in range(world)]
        else:
            ranges = nnz_balanced_ranges(g.rowptr, world)
        lo, hi = ranges[rank]
        shard = g.row_slice(lo, hi)
        rp, ci = shard.host_arrays()
        rp, ci = rp - rp[0], ci[rp[0]:]      # drop the 32-arc alignment pad of the slice
        cur = X0
        for _ in range(levels):
            nxt = torch.full((g.n, d), float('nan'), dtype=torch.float64)
            _, M = oracle.aggregate_csr(rp, ci, cur.numpy())
            nxt[lo:hi] = torch.from_numpy(M)
            exchange_rows(nxt, ranges, rank, dist)
            cur = nxt
        ret[rank] = cur.numpy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('equal_rows', [False, True, 'cost'])
def test_two_rank_recursion_equals_single_process(equal_rows):
    world, levels = (3, 3) if equal_rows == 'cost' else (2, 3)
    port = _free_port()
    manager = mp.Manager()
    ret = manager.dict()
    mp.spawn(_worker, args=(world, port, levels, equal_rows, ret), nprocs=world, join=True)
    g = barabasi_albert_csr(3000, 6, seed=3, device='cpu')
    rp, ci = g.host_arrays()
    cur = torch.rand(g.n, 5, generator=torch.Generator().manual_seed(0),
                     dtype=torch.float64).numpy()
    for _ in range(levels):
        _, cur = oracle.aggregate_csr(rp, ci, cur)
    for rank in range(world):
        np.testing.assert_allclose(ret[rank], cur, rtol=1e-13, atol=0)


def _grid_worker(rank, world, port, levels, ret):
    """C = 2 column groups x R = 2 node ranges on 4 gloo ranks: the exchange of a column group
    runs in its own process group and never sees the other group's columns."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        C, R = 2, 2
        c, r = rank % C, rank // C
        g = barabasi_albert_csr(2000, 5, seed=7, device='cpu')
        d = 8
        c0, c1 = column_groups(d, C)[c]
        X0 = torch.rand(g.n, d, generator=torch.Generator().manual_seed(0), dtype=torch.float64)
        groups = [dist.new_group([q * C + cc for q in range(R)]) for cc in range(C)]
        members = [q * C + c for q in range(R)]
        ranges = cost_balanced_ranges(g.rowptr, R, 10.0)
        lo, hi = ranges[r]
        shard = g.row_slice(lo, hi)
        rp, ci = shard.host_arrays()
        rp, ci = rp - rp[0], ci[rp[0]:]
        cur = X0[:, c0:c1].contiguous()
        for _ in range(levels):
            nxt = torch.full((g.n, c1 - c0), float('nan'), dtype=torch.float64)
            _, M = oracle.aggregate_csr(rp, ci, cur.numpy())
            nxt[lo:hi] = torch.from_numpy(M)
            exchange_rows(nxt, ranges, r, groups[c], members)
            cur = nxt
        ret[rank] = (c0, c1, cur.numpy())
    finally:
        dist.destroy_process_group()


def test_column_groups_times_node_ranges_on_four_ranks():
    world, levels = 4, 3
    port = _free_port()
    manager = mp.Manager()
    ret = manager.dict()
    mp.spawn(_grid_worker, args=(world, port, levels, ret), nprocs=world, join=True)
    g = barabasi_albert_csr(2000, 5, seed=7, device='cpu')
    rp, ci = g.host_arrays()
    cur = torch.rand(g.n, 8, generator=torch.Generator().manual_seed(0),
                     dtype=torch.float64).numpy()
    for _ in range(levels):
        _, cur = oracle.aggregate_csr(rp, ci, cur)
    for rank in range(world):
        c0, c1, got = ret[rank]
        np.testing.assert_allclose(got, cur[:, c0:c1], rtol=1e-13, atol=0)


# ---- path B: row-sharded NMF loop (roles/sharded.py) ------------------------------------------
class _OracleNmfBackend:
    """The three compute steps of RowShardedNmf done by the float64 oracle (CPU tensors)."""

    def local_iteration(self, X, W, H, use_tf32=True):
        Xn, Hn = X.numpy(), H.numpy()
        Wn = nmf_oracle.update_w(Xn, W.numpy(), Hn)
        W.copy_(torch.from_numpy(Wn))
        return torch.from_numpy(np.concatenate([(Wn.T @ Xn).ravel(), (Wn.T @ Wn).ravel()]))

    def update_h(self, sums, H):
        r, f = H.shape
        numer = sums[:r * f].numpy().reshape(r, f)
        denom = sums[r * f:].numpy().reshape(r, r) @ H.numpy()
        denom[denom == 0] = nmf_oracle.EPSILON
        H.copy_(torch.from_numpy(H.numpy() * (numer / denom)))

    def error_sq(self, X, W, H, use_tf32=True):
        return nmf_oracle.frobenius_error(X.numpy(), W.numpy(), H.numpy()) ** 2


def _nmf_problem():
    rng = np.random.RandomState(4)
    X = (rng.rand(700, 5) ** 2) @ rng.rand(5, 24) + 0.05 * rng.rand(700, 24)
    return X, rng.rand(700, 4) + 0.1, rng.rand(4, 24) + 0.1


def _nmf_worker(rank, world, port, ret):
    from graphrole_b200.roles.sharded import RowShardedNmf, row_shard
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        X, W0, H0 = _nmf_problem()
        lo, hi = row_shard(X.shape[0], world, rank, align=64)
        Xl, Wl = torch.from_numpy(X[lo:hi].copy()), torch.from_numpy(W0[lo:hi].copy())
        H = torch.from_numpy(H0.copy())
        solver = RowShardedNmf(hi - lo, X.shape[1], 4, backend=_OracleNmfBackend())
        n_iter, err = solver.fit(Xl, Wl, H, tol=2e-3)
        ret[rank] = (lo, hi, Wl.numpy(), H.numpy(), n_iter, err)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_row_sharded_nmf_loop_equals_single_process(world):
    """RowShardedNmf.fit on gloo ranks (oracle compute): same factors, same stopping iteration
    and same error as the single-process float64 loop -- the all-reduce of [W^T X | W^T W] and
    of the squared residual is all that crosses the ranks; every rank ends with the same H."""
    from graphrole_b200.roles.sharded import row_shard
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_nmf_worker, args=(world, port, ret), nprocs=world, join=True)
    X, W0, H0 = _nmf_problem()
    W, H, n_iter = nmf_oracle.fit_multiplicative_update(X, W0, H0, tol=2e-3)
    err = nmf_oracle.frobenius_error(X, W, H)
    covered = 0
    for rank in range(world):
        lo, hi, Wl, Hl, it, e = ret[rank]
        assert (lo, hi) == row_shard(X.shape[0], world, rank, align=64)
        covered += hi - lo
        assert it == n_iter and it % 10 == 0 and it < 200
        np.testing.assert_allclose(Wl, W[lo:hi], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(Hl, H, rtol=1e-9, atol=1e-12)
        np.testing.assert_array_equal(Hl, ret[0][3])
        assert e == pytest.approx(err, rel=1e-10)
    assert covered == X.shape[0]


def test_row_shard_covers_all_rows_in_whole_blocks():
    from graphrole_b200.roles.sharded import row_shard
    for n, world in [(10_000_000, 8), (1000, 3), (129, 2), (5, 4)]:
        spans = [row_shard(n, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(lo % 128 == 0 for lo, _ in spans if lo < n)


def _empty_shard_worker(rank, world, port, ret):
    from graphrole_b200.roles.sharded import RowShardedNmf, row_shard
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lo, hi = row_shard(130, world, rank)          # 128-row blocks: the third rank gets nothing
        try:
            RowShardedNmf(hi - lo, 8, 2, backend=_OracleNmfBackend())
            ret[rank] = 'constructed'
        except ValueError as exc:
            ret[rank] = 'ValueError: ' + str(exc)
    finally:
        dist.destroy_process_group()


def test_row_sharded_nmf_refuses_an_empty_shard_on_every_rank():
    """A rank without rows cannot create its workspaces; the refusal is voted on, so ALL ranks
    raise instead of the healthy ones waiting in the first all-reduce."""
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_empty_shard_worker, args=(3, port, ret), nprocs=3, join=True)
    assert all(ret[r].startswith('ValueError') for r in range(3)), dict(ret)
