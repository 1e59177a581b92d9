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
from graphrole_b200.shard import cost_balanced_ranges, exchange_rows, nnz_balanced_ranges
from oracle import refex_oracle as oracle


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
