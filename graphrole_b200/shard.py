"""Node-range sharding of the ReFeX recursion across the GPUs of one box.

Path A shards by output row: rank g owns the contiguous node range [lo_g, hi_g) (boundaries
balance the ranks' level time: arcs gathered, and for the fused exchange also rows stored to the
peers -- power-law graphs put the hubs at the front), holds that slice of the CSR, a full replica of the current level's input matrix, and
produces its rows of the next level.  The one exchange step per level is an all-gather of the
block the recursion continues on.  Primary form: the exchange is fused
into the gather kernel -- every mean row is stored into all ranks' replicas of the next input
matrix through NVLink-mapped peer pointers, and a flag barrier in the same mapped memory
separates the levels (PeerReplicas).  Alternative form: every rank writes its slice and the
slices are exchanged in place by a torch.distributed all-gather (exchange_rows).
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


def cost_balanced_ranges(rowptr: torch.Tensor, world: int, row_cost: float) -> List[Tuple[int, int]]:
    """Contiguous row ranges that minimise the slowest rank when a rank's level time is
    max(arcs gathered, row_cost * rows produced) -- the fused exchange: gathering an arc costs
    d*4 bytes of HBM traffic, producing a row costs d*4 bytes of NVLink stores to each of the
    world-1 peers, and the two overlap inside one kernel.  row_cost is the price of a row in
    arcs: (HBM bandwidth / NVLink store bandwidth) * (world - 1).  Measured on C3 with
    arc-balanced ranges the LAST rank (millions of low-degree rows) was bound by its stores:
    4 GPUs 5.4 ms per level for 3.4 GB, 8 GPUs 7.0 ms for 4.2 GB, i.e. ~0.6 TB/s against
    ~6.2 TB/s of gather traffic -- hence the default ratio of 10.
    Binary search on the per-rank budget, greedy assignment from row 0; row_cost == 0 gives
    (nearly) the arc-balanced ranges."""
    n = rowptr.numel() - 1
    if world == 1 or n == 0:
        return [(0, n)] + [(n, n)] * (world - 1)
    rp = rowptr.detach().to('cpu')
    nnz = int(rp[-1])
    base = int(rp[0])

    def assign(budget: float):
        bounds, lo = [0], 0
        for _ in range(world):
            if lo >= n:
                bounds.append(n)
                continue
            # furthest hi with arcs(lo, hi) <= budget and (hi - lo) * row_cost <= budget
            hi_arcs = int(torch.searchsorted(rp, torch.tensor(int(rp[lo]) + int(budget)),
                                             right=True)) - 1
            hi_rows = n if row_cost <= 0 else lo + int(budget / row_cost)
            hi = max(lo + 1, min(hi_arcs, hi_rows, n))       # always make progress
            bounds.append(hi)
            lo = hi
        return bounds

    lo_b, hi_b = 1.0, float(max(nnz - base, 1) + row_cost * n + 1)
    for _ in range(60):
        mid = 0.5 * (lo_b + hi_b)
        if assign(mid)[-1] >= n:
            hi_b = mid
        else:
            lo_b = mid
        if hi_b - lo_b <= 1.0:
            break
    bounds = assign(hi_b)
    bounds[-1] = n
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


class PeerReplicas:
    """Two full [n, d] fp32 replicas of the level input plus the barrier flags, in ONE
    peer-mappable buffer per rank (gr_peer_alloc), with every other rank's buffer mapped into
    this process.  Layout: replica 0 | replica 1 | flag words (256-byte aligned offsets)."""

    def __init__(self, n: int, d: int, world: int, rank: int, device: torch.device):
        import torch.distributed as dist
        from graphrole_b200 import _native
        self.n, self.d, self.world, self.rank = n, d, world, rank
        self.replica_bytes = (n * d * 4 + 255) // 256 * 256
        self.flag_offset = 2 * self.replica_bytes
        nbytes = self.flag_offset + 8 * _native.peer_flag_words()
        dev = device.index if device.index is not None else torch.cuda.current_device()
        self.own = _native.PeerBuffer(nbytes, dev)
        # the flag words live right behind the replicas
        self.flags = torch.as_tensor(self.own, device=device)[self.flag_offset:].view(torch.int64)
        self.flags.zero_()
        torch.cuda.synchronize(device)
        tokens = [None] * world
        dist.all_gather_object(tokens, self.own.handle)     # also orders the zeroing
        self.buffers = []
        for q in range(world):
            self.buffers.append(self.own if q == rank
                                else _native.PeerBuffer.open(tokens[q], nbytes, dev))
        self.epoch = 0
        flat = torch.as_tensor(self.own, device=device)
        self.replicas = [flat[i * self.replica_bytes: i * self.replica_bytes + n * d * 4]
                         .view(torch.float32).view(n, d) for i in range(2)]

    def replica_ptrs(self, which: int):
        return [b.ptr + which * self.replica_bytes for b in self.buffers]

    def flag_ptrs(self):
        return [b.ptr + self.flag_offset for b in self.buffers]

    def barrier(self, stream=None):
        from graphrole_b200 import _native
        self.epoch += 1
        _native.peer_barrier(self.flag_ptrs(), self.rank, self.epoch, stream=stream)

    def timed_out(self) -> int:
        from graphrole_b200 import _native
        return _native.peer_barrier_timed_out(self.own.ptr + self.flag_offset)

    def close(self):
        """Unmap the peers' buffers, then (after everyone has) release the own one."""
        import torch.distributed as dist
        torch.cuda.synchronize()
        for q, b in enumerate(self.buffers):
            if q != self.rank:
                b.close()
        if dist.is_initialized():
            dist.barrier()
        self.replicas = None
        self.flags = None
        self.own.close()


class ShardedRefex:
    """Runs `levels` recursion levels (schedule alpha: recurse on the mean block) on this
    rank's node range; world == 1 is the plain single-GPU path with no exchange.

    exchange (world > 1):
      'peer'  (default) the gather kernel stores every mean row into all ranks' replicas of the
              next input matrix over NVLink-mapped pointers (gr_refex_aggregate_bcast_f32) and
              a flag barrier in the same mapped memory separates the levels -- no collective
              library call on the data path;
      'nccl'  the kernel writes the own slice and a torch.distributed all-gather (or one
              broadcast per owner for uneven ranges) exchanges the slices in place.
    If the CUDA IPC mapping cannot be set up on every rank, all ranks drop to 'nccl' together
    and `exchange_note` says why."""

    def __init__(self, graph: CSRGraph, d: int, world: int = 1, rank: int = 0, group=None,
                 exchange: Optional[str] = None):
        self.d = d
        self.world = world
        self.rank = rank
        self.group = group
        self.n = graph.n
        self.exchange = 'none'
        self.exchange_note = ''
        self.peers = None
        device = graph.rowptr.device
        self.balance = 'single range'
        if world == 1:
            self.ranges = [(0, graph.n)]
            self.handle = graph.handle(device)
            self.local_rows, self.local_nnz = graph.n, graph.nnz
            self.out = [torch.empty((graph.n, 2 * d), dtype=torch.float32, device=device)
                        for _ in range(2)]
            return
        import os
        exchange = exchange or os.environ.get('GR_SHARD_EXCHANGE', 'peer')
        if exchange not in ('peer', 'nccl'):
            raise ValueError("exchange must be 'peer' or 'nccl'")
        # fused exchange: a rank's stores to its world-1 peers overlap its gathers, so ranges
        # balance max(arcs, row_cost * rows); the all-gather form balances arcs alone
        ratio = float(os.environ.get('GR_SHARD_HBM_NVLINK_RATIO', '10'))
        if exchange == 'peer' and ratio > 0 and os.environ.get('GR_SHARD_BALANCE', 'cost') == 'cost':
            self.balance = f'max(arcs, {ratio * (world - 1):g} x rows)-balanced ranges'
            self.ranges = cost_balanced_ranges(graph.rowptr, world, ratio * (world - 1))
        else:
            self.balance = 'arc-balanced ranges'
            self.ranges = nnz_balanced_ranges(graph.rowptr, world)
        lo, hi = self.ranges[rank]
        self.shard = graph.row_slice(lo, hi)
        self.handle = self.shard.handle(device)
        self.local_rows, self.local_nnz = hi - lo, self.shard.nnz
        self.sums = torch.empty((hi - lo, d), dtype=torch.float32, device=device)
        if exchange == 'peer':
            import torch.distributed as dist
            ok, why = 1, ''
            try:
                self.peers = PeerReplicas(graph.n, d, world, rank, device)
            except Exception as exc:      # every rank must take the same path: vote below
                ok, why = 0, repr(exc)
            vote = torch.tensor([ok], device=device, dtype=torch.int32)
            dist.all_reduce(vote, op=dist.ReduceOp.MIN)
            if int(vote.item()) == 0:
                if self.peers is not None:
                    self.peers = None
                exchange = 'nccl'
                self.exchange_note = 'peer mapping failed on some rank, using nccl: ' + why
        self.exchange = exchange
        if exchange == 'nccl':
            self.full = [torch.empty((graph.n, d), dtype=torch.float32, device=device)
                         for _ in range(2)]

    def run_levels(self, X0: torch.Tensor, levels: int, events: Optional[list] = None):
        """Returns the last level's (sum rows, mean rows) held by this rank."""
        d = self.d
        cur = X0
        last = None
        if self.exchange == 'peer':
            # the own replica 0 is the level-0 input (every rank holds the same X0)
            self.peers.replicas[0].copy_(X0)
            cur = self.peers.replicas[0]
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
            elif self.exchange == 'peer':
                lo, hi = self.ranges[self.rank]
                which = (level + 1) & 1
                self.handle.aggregate_bcast(cur, self.sums, self.peers.replica_ptrs(which), d, lo)
                if events is not None:
                    e1.record()
                    events.append((e0, e1))
                self.peers.barrier()
                cur = self.peers.replicas[which]
                last = (self.sums, cur[lo:hi])
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

    def close(self):
        if self.peers is not None:
            if self.peers.timed_out():
                raise RuntimeError(f'peer barrier timed out at epoch {self.peers.timed_out()}')
            self.peers.close()
            self.peers = None
