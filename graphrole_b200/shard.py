"""Node-range sharding of the ReFeX recursion across the GPUs of one box.

Path A shards by output row: rank g owns the contiguous node range [lo_g, hi_g) (boundaries
chosen so every rank traverses the same number of arcs -- power-law graphs put the hubs at the
front), holds that slice of the CSR, a full replica of the current level's input matrix, and
produces its rows of the next level.  The one exchange step per level is an all-gather of the
block the recursion continues on: every rank writes its mean rows
straight into its slice of the next full input matrix and the slices are exchanged in place.
"""
from typing import List, Optional, Tuple

import torch

from graphrole_b200.graph.csr import CSRGraph


def nnz_balanced_ranges(rowptr: torch.Tensor, world: int) -> List[Tuple[int, int]]:
    """Contiguous row ranges with (nearly) equal arc counts; every row belongs to one range."""
    n = rowptr.numel() - 1
    nnz = int(rowptr[-1])
    targets = torch.tensor([nnz * k // world for k in range(1, world)], dtype=rowptr.dtype,
                           device=rowptr.device)
    cuts = torch.searchsorted(rowptr, targets, right=False).clamp_(0, n).tolist()
    bounds = [0] + [int(c) for c in cuts] + [n]
    for i in range(1, len(bounds)):          # keep monotone (empty ranges are allowed)
        bounds[i] = max(bounds[i], bounds[i - 1])
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


def exchange_rows(full: torch.Tensor, ranges: List[Tuple[int, int]], rank: int, group) -> None:
    """In-place all-gather of row slices: on entry rank g has filled full[lo_g:hi_g]; on exit
    every rank holds all rows.  `group` is the torch.distributed module/process group owner."""
    import torch.distributed as dist
    views = [full[lo:hi] for lo, hi in ranges]
    sizes = {hi - lo for lo, hi in ranges}
    if len(sizes) == 1 and full.is_contiguous():
        dist.all_gather_into_tensor(full, views[rank])
        return
    # uneven slices: one broadcast per owner, issued back to back on the collective stream
    works = [dist.broadcast(v, src=g, async_op=True) for g, v in enumerate(views) if v.numel()]
    for w in works:
        w.wait()


class ShardedRefex:
    """Runs `levels` recursion levels (schedule alpha: recurse on the mean block) on this
    rank's node range; world == 1 is the plain single-GPU path with no exchange."""

    def __init__(self, graph: CSRGraph, d: int, world: int = 1, rank: int = 0, group=None):
        self.d = d
        self.world = world
        self.rank = rank
        self.group = group
        self.n = graph.n
        device = graph.rowptr.device
        if world == 1:
            self.ranges = [(0, graph.n)]
            self.handle = graph.handle(device)
            self.local_rows, self.local_nnz = graph.n, graph.nnz
            self.out = [torch.empty((graph.n, 2 * d), dtype=torch.float32, device=device)
                        for _ in range(2)]
        else:
            self.ranges = nnz_balanced_ranges(graph.rowptr, world)
            lo, hi = self.ranges[rank]
            self.shard = graph.row_slice(lo, hi)
            self.handle = self.shard.handle(device)
            self.local_rows, self.local_nnz = hi - lo, self.shard.nnz
            self.sums = torch.empty((hi - lo, d), dtype=torch.float32, device=device)
            self.full = [torch.empty((graph.n, d), dtype=torch.float32, device=device)
                         for _ in range(2)]

    def run_levels(self, X0: torch.Tensor, levels: int, events: Optional[list] = None):
        """Returns the last level's (sum rows, mean rows) held by this rank."""
        d = self.d
        cur = X0
        last = None
        for level in range(levels):
            if events is not None:
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            if self.world == 1:
                out = self.handle.aggregate(cur, out=self.out[level & 1])
                if events is not None:
                    e1.record()
                    events.append((e0, e1))
                cur = out[:, d:]
                last = (out[:, :d], out[:, d:])
            else:
                lo, hi = self.ranges[self.rank]
                nxt = self.full[level & 1]
                self.handle.aggregate_into(cur, self.sums, nxt[lo:hi])
                if events is not None:
                    e1.record()
                    events.append((e0, e1))
                exchange_rows(nxt, self.ranges, self.rank, self.group)
                cur = nxt
                last = (self.sums, nxt[lo:hi])
        return last
