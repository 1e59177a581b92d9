"""Sharding of the ReFeX recursion across the GPUs of one box.

The ranks form a grid of C column groups x R node ranges (world = C * R):

* node ranges (R): rank (c, r) owns the contiguous node range [lo_r, hi_r), holds that slice of
  the CSR, a full replica of its column group's level input and produces its rows of the next
  level.  The one exchange step per level is an all-gather, inside the column group, of the block
  the recursion continues on.  Primary form: the exchange is fused into the gather kernel -- every
  mean row is stored into all group members' replicas of the next input through NVLink-mapped peer
  pointers, and a flag barrier in the same mapped memory separates the levels (PeerReplicas).
  Alternative form: every rank writes its slice and the slices are exchanged in place by a
  torch.distributed all-gather (exchange_rows).
* column groups (C): aggregation is column-separable (S[:, j] = A X[:, j]), so a group that owns
  d / C of the feature columns needs nothing from the other groups, ever.  Splitting columns
  divides the exchange volume per GPU by C (2.24 GB per level for 8 node ranges of C3, 0.96 GB for
  2 x 4) at the price of narrower gathers: measured on one B200 (profiles/r2_colsplit.txt), a C3
  level over all rows takes 17.8 / 10.4 / 7.8 / 6.0 ms at d = 64 / 32 / 16 / 8 -- random gathers
  below 256 bytes are bound by DRAM row activations, not bytes -- so columns are split only where
  the exchange would otherwise bound the level (8 GPUs: 2 x 4).

Ranges start from a cost model (arcs gathered vs rows stored to the peers) and are then corrected
by MEASURED feedback: `autobalance` times every rank's kernel on the real input, all-gathers the
times and re-cuts the ranges so that the predicted time per rank is equal.
"""
import os
from typing import List, Optional, Tuple

import torch

from graphrole_b200.graph.csr import CSRGraph


def bind_host_thread_to_gpu(device_index: int) -> str:
    """Pin the calling thread to the CPU cores next to GPU `device_index` (NVML's ideal CPU
    affinity), so that the pinned host buffers it allocates afterwards land on that GPU's NUMA
    node.  One process per GPU all first-touching their staging buffers on the same node is what
    limits the host-buffer path on an 8-GPU box: measured 14.5 GB/s of D2H per GPU at 8 ranks
    against 50 GB/s for one.  Returns a note for the bench line; never raises."""
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        handle = None
        uuid = getattr(props, 'uuid', None)
        if uuid is not None:
            try:
                handle = pynvml.nvmlDeviceGetHandleByUUID(f'GPU-{uuid}'.encode())
            except Exception:
                handle = None
        if handle is None:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        cpus = sorted(os.sched_getaffinity(0))
        return f'host thread bound to {len(cpus)} CPUs next to the GPU ({cpus[0]}..{cpus[-1]})'
    except Exception as exc:          # no NVML, no permission, a container without cpusets ...
        return f'host thread not bound ({exc!r})'


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
    peers, and the two overlap inside one kernel.  row_cost is the price of a row in arcs.
    Binary search on the per-rank budget, greedy assignment from row 0; row_cost == 0 gives
    (nearly) the arc-balanced ranges.  This is only the STARTING point: see time_balanced_ranges."""
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


def time_balanced_ranges(rowptr: torch.Tensor, ranges: List[Tuple[int, int]],
                         times_ms: List[float], row_cost: float = 8.0,
                         damping: float = 1.0) -> List[Tuple[int, int]]:
    """Re-cut contiguous ranges from MEASURED per-range times.  Inside an old range the time is
    spread over its rows in proportion to (arcs of the row + row_cost): that gives a cumulative
    time T(row) over the whole graph, and the new boundaries are where T crosses k/R of its
    total.  `damping` < 1 moves the boundaries only part of the way (measurements are noisy and
    the speed inside a range is not really uniform).  Pure host arithmetic on rowptr."""
    world = len(ranges)
    n = rowptr.numel() - 1
    if world == 1 or n == 0:
        return list(ranges)
    rp = rowptr.detach().to('cpu', torch.float64)
    unit = (rp[1:] - rp[:-1]) + float(row_cost)                # work units of every row
    density = torch.zeros(n, dtype=torch.float64)
    for (lo, hi), t in zip(ranges, times_ms):
        if hi > lo:
            w = float(unit[lo:hi].sum())
            density[lo:hi] = max(float(t), 1e-6) / max(w, 1e-9)
    cum = torch.cumsum(unit * density, 0)                       # T(row + 1)
    total = float(cum[-1])
    targets = torch.tensor([total * k / world for k in range(1, world)], dtype=torch.float64)
    cuts = (torch.searchsorted(cum, targets, right=False) + 1).clamp_(0, n).tolist()
    old = [r[1] for r in ranges[:-1]]
    cuts = [int(round(o + damping * (c - o))) for o, c in zip(old, cuts)]
    bounds = [0] + cuts + [n]
    for i in range(1, len(bounds)):
        bounds[i] = min(max(bounds[i], bounds[i - 1]), n)
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


def column_groups(d: int, groups: int) -> List[Tuple[int, int]]:
    """Split d feature columns into `groups` contiguous blocks whose boundaries are multiples of
    four columns (16-byte vector loads) whenever d allows."""
    if groups < 1 or groups > max(d, 1):
        raise ValueError(f'cannot split {d} columns into {groups} groups')
    quad = 4 if d % 4 == 0 and d // 4 >= groups else 1
    units = d // quad
    bounds = [quad * (units * k // groups) for k in range(groups + 1)]
    return [(bounds[k], bounds[k + 1]) for k in range(groups)]


def default_column_groups(world: int, d: int) -> int:
    """Column groups used when the caller does not say: split columns only where the node-range
    exchange would bound the level (8 GPUs and at least 32 columns per group)."""
    env = os.environ.get('GR_SHARD_COL_GROUPS')
    if env:
        c = int(env)
        if c < 1 or world % c or c > d:
            raise ValueError(f'GR_SHARD_COL_GROUPS={c} does not divide world={world} / d={d}')
        return c
    if world >= 8 and world % 2 == 0 and d % 8 == 0 and d // 2 >= 32:
        return 2
    return 1


def exchange_rows(full: torch.Tensor, ranges: List[Tuple[int, int]], rank: int, group=None,
                  ranks: Optional[List[int]] = None) -> None:
    """In-place all-gather of row slices: on entry member `rank` (index into `ranges`) has filled
    full[lo:hi]; on exit every member holds all rows.  `group` is a torch.distributed process
    group (None = the default group) whose global ranks are `ranks` (default 0..len-1)."""
    import torch.distributed as dist
    group = None if group is dist else group        # older callers pass the module itself
    views = [full[lo:hi] for lo, hi in ranges]
    sizes = {hi - lo for lo, hi in ranges}
    if len(sizes) == 1 and full.is_contiguous():
        dist.all_gather_into_tensor(full, views[rank], group=group)
        return
    # uneven slices: one broadcast per owner, issued back to back on the collective stream
    ranks = list(range(len(ranges))) if ranks is None else ranks
    works = [dist.broadcast(v, src=ranks[g], group=group, async_op=True)
             for g, v in enumerate(views) if v.numel()]
    for w in works:
        w.wait()


class PeerReplicas:
    """Two full [n, d] fp32 replicas of the level input plus the barrier flags, in ONE
    peer-mappable buffer per rank (gr_peer_alloc), with the buffers of the other members of the
    exchange group mapped into this process.  Layout: replica 0 | replica 1 | flag words
    (256-byte aligned offsets).

    Set-up is split so that a failure on one rank can never leave the ranks in different
    collectives: allocate() is local, tokens are exchanged only after every rank voted that its
    allocation worked, and connect() is followed by a second vote (ShardedRefex does the voting).
    """

    def __init__(self, n: int, d: int, members: List[int], rank: int, device: torch.device):
        self.n, self.d = n, d
        self.members = list(members)            # global ranks of the exchange group
        self.index = self.members.index(rank)   # my position in it
        self.device = device
        self.replica_bytes = (n * d * 4 + 255) // 256 * 256
        self.flag_offset = 2 * self.replica_bytes
        self.own = None
        self.buffers = []
        self.replicas = None
        self.flags = None
        self.epoch = 0

    def allocate(self):
        from graphrole_b200 import _native
        nbytes = self.flag_offset + 8 * _native.peer_flag_words()
        dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.nbytes, self.dev = nbytes, dev
        self.own = _native.PeerBuffer(nbytes, dev)
        flat = torch.as_tensor(self.own, device=self.device)
        # the flag words live right behind the replicas
        self.flags = flat[self.flag_offset:].view(torch.int64)
        self.flags.zero_()
        torch.cuda.synchronize(self.device)
        self.replicas = [flat[i * self.replica_bytes: i * self.replica_bytes + self.n * self.d * 4]
                         .view(torch.float32).view(self.n, self.d) for i in range(2)]
        return self.own.handle

    def connect(self, tokens_by_rank):
        """tokens_by_rank[q] = rank q's IPC token (all ranks of the job)."""
        from graphrole_b200 import _native
        self.buffers = []
        try:
            for q in self.members:
                self.buffers.append(self.own if q == self.members[self.index] else
                                    _native.PeerBuffer.open(tokens_by_rank[q], self.nbytes,
                                                            self.dev))
        except Exception:
            self.release(collective=False)
            raise

    def replica_ptrs(self, which: int):
        return [b.ptr + which * self.replica_bytes for b in self.buffers]

    def flag_ptrs(self):
        return [b.ptr + self.flag_offset for b in self.buffers]

    def barrier(self, stream=None):
        from graphrole_b200 import _native
        self.epoch += 1
        _native.peer_barrier(self.flag_ptrs(), self.index, self.epoch, stream=stream)

    def timed_out(self) -> int:
        from graphrole_b200 import _native
        return _native.peer_barrier_timed_out(self.own.ptr + self.flag_offset)

    def release(self, collective: bool = True):
        """Unmap the peers' buffers, then (after everyone has) release the own one."""
        import torch.distributed as dist
        torch.cuda.synchronize()
        for b in self.buffers:
            if b is not self.own:
                b.close()
        self.buffers = []
        if collective and dist.is_initialized():
            dist.barrier()
        self.replicas = None
        self.flags = None
        if self.own is not None:
            self.own.close()
            self.own = None

    close = release


class ShardedRefex:
    """Runs `levels` recursion levels (schedule alpha: recurse on the mean block) on this rank's
    share of the work: column group c = rank % C, node range r = rank // C.  world == 1 is the
    plain single-GPU path with no exchange.

    exchange (R > 1):
      'peer'  (default) the gather kernel stores every mean row into all group members' replicas
              of the next input over NVLink-mapped pointers (gr_refex_aggregate_bcast_f32) and a
              flag barrier in the same mapped memory separates the levels -- no collective
              library call on the data path;
      'nccl'  the kernel writes the own slice and a torch.distributed all-gather (or one
              broadcast per owner for uneven ranges) exchanges the slices in place.
    If the CUDA IPC mapping cannot be set up on every rank, all ranks drop to 'nccl' together
    and `exchange_note` says why."""

    def __init__(self, graph: CSRGraph, d: int, world: int = 1, rank: int = 0, group=None,
                 exchange: Optional[str] = None, col_groups: Optional[int] = None):
        self.graph = graph
        self.d_total = d
        self.world, self.rank = world, rank
        self.n = graph.n
        self.exchange = 'none'
        self.exchange_note = ''
        self.peers = None
        self.device = graph.rowptr.device if graph.rowptr.is_cuda else \
            torch.device('cuda', torch.cuda.current_device())
        C = col_groups if col_groups else default_column_groups(world, d)
        if world % C or C > d:
            raise ValueError(f'{C} column groups do not divide world={world} / d={d}')
        self.C, self.R = C, world // C
        self.c, self.r = rank % C, rank // C
        self.col_lo, self.col_hi = column_groups(d, C)[self.c]
        self.d = self.col_hi - self.col_lo          # columns this rank aggregates
        self.members = [q * C + self.c for q in range(self.R)]   # my exchange group
        self.group = None
        self.balance = 'single range'
        self.balance_history = []
        self.out = None
        if self.R == 1:
            self.ranges = [(0, graph.n)]
            self._build_shard()
            self.out = [torch.empty((graph.n, 2 * self.d), dtype=torch.float32,
                                    device=self.device) for _ in range(2)]
            return

        import torch.distributed as dist
        exchange = exchange or os.environ.get('GR_SHARD_EXCHANGE', 'peer')
        if exchange not in ('peer', 'nccl'):
            raise ValueError("exchange must be 'peer' or 'nccl'")
        # starting ranges: the fused exchange balances max(arcs, row_cost * rows) -- a rank's
        # stores to its R - 1 peers overlap its gathers; the all-gather form balances arcs alone
        ratio = float(os.environ.get('GR_SHARD_HBM_NVLINK_RATIO', '10'))
        if exchange == 'peer' and ratio > 0 and os.environ.get('GR_SHARD_BALANCE', 'cost') == 'cost':
            self.balance = f'max(arcs, {ratio * (self.R - 1):g} x rows)-balanced ranges'
            self.ranges = cost_balanced_ranges(graph.rowptr, self.R, ratio * (self.R - 1))
        else:
            self.balance = 'arc-balanced ranges'
            self.ranges = nnz_balanced_ranges(graph.rowptr, self.R)
        self._build_shard()

        if exchange == 'peer':
            # (1) local allocation, (2) vote, (3) token exchange, (4) mapping, (5) vote: every rank
            # runs the same sequence of collectives whatever fails where
            ok, why, token = 1, '', None
            peers = PeerReplicas(graph.n, self.d, self.members, rank, self.device)
            try:
                token = peers.allocate()
            except Exception as exc:
                ok, why = 0, repr(exc)
            if self._all_ok(ok):
                tokens = [None] * world
                dist.all_gather_object(tokens, token)
                try:
                    peers.connect(tokens)
                except Exception as exc:
                    ok, why = 0, repr(exc)
                if not self._all_ok(ok):
                    peers.release(collective=True)
                    exchange = 'nccl'
            else:
                if peers.own is not None:
                    peers.release(collective=False)
                exchange = 'nccl'
            if exchange == 'peer':
                self.peers = peers
            else:
                self.exchange_note = ('peer mapping failed on some rank, using nccl'
                                      + (': ' + why if why else ''))
        self.exchange = exchange
        if exchange == 'nccl':
            # one process group per column group; every rank creates all of them
            for c in range(C):
                grp = dist.new_group([q * C + c for q in range(self.R)]) if C > 1 else None
                if c == self.c:
                    self.group = grp
            self.full = [torch.empty((graph.n, self.d), dtype=torch.float32, device=self.device)
                         for _ in range(2)]

    # ---- set-up helpers ----------------------------------------------------------------------
    def _all_ok(self, ok: int) -> bool:
        import torch.distributed as dist
        vote = torch.tensor([ok], device=self.device, dtype=torch.int32)
        dist.all_reduce(vote, op=dist.ReduceOp.MIN)
        return int(vote.item()) == 1

    def _build_shard(self):
        lo, hi = self.ranges[self.r]
        if self.R == 1:
            self.shard = self.graph
            self.handle = self.graph.handle(self.device)
        else:
            self.shard = self.graph.row_slice(lo, hi)
            self.handle = self.shard.handle(self.device)
        if self.d * 4 != 256:
            # the hot-row budget is sized for 256-byte rows: narrower rows pin more of them
            self.handle.tune_hot_rows(self.d * 4)
        self.local_rows = hi - lo
        self.local_nnz = self.shard.nnz
        if self.R > 1:
            self.sums = torch.empty((hi - lo, self.d), dtype=torch.float32, device=self.device)

    # ---- measured balancing ---------------------------------------------------------------
    def autobalance(self, X0: torch.Tensor, levels: int, rounds: int = 4,
                    tolerance: float = 1.02) -> List[dict]:
        """Time this rank's kernel on the real input, all-gather the times, re-cut the node
        ranges so that the predicted per-rank time is equal; stop when max / mean <= tolerance
        or after `rounds` corrections.  Every rank derives the same ranges from the same
        gathered numbers.  Returns the history [{ranges, ms per range}]."""
        if self.R == 1:
            return []
        import torch.distributed as dist
        best = None                                  # (slowest range's ms, ranges)
        for it in range(rounds + 1):
            events = []
            self.run_levels(X0, levels, events)            # first pass also warms everything up
            events = []
            self.run_levels(X0, levels, events)
            torch.cuda.synchronize()
            mine = sum(ev[0].elapsed_time(ev[1]) for ev in events) / max(levels, 1)
            t = torch.zeros(self.world, device=self.device, dtype=torch.float64)
            t[self.rank] = mine
            dist.all_reduce(t)
            per_rank = t.tolist()
            # ranks with the same node range (different column groups) do the same work
            per_range = [max(per_rank[q * self.C + c] for c in range(self.C))
                         for q in range(self.R)]
            self.balance_history.append({'ranges': list(self.ranges),
                                         'kernel_ms': [round(x, 4) for x in per_range]})
            if best is None or max(per_range) < best[0]:
                best = (max(per_range), list(self.ranges))
            mean = sum(per_range) / len(per_range)
            if it == rounds or max(per_range) <= tolerance * mean:
                break
            new = time_balanced_ranges(self.graph.rowptr, self.ranges, per_range,
                                       damping=1.0 if it == 0 else 0.7)
            if new == self.ranges:
                break
            self.ranges = new
            self._build_shard()
            self.balance = f'measured-time-balanced ranges ({it + 1} correction(s))'
        if best is not None and best[1] != self.ranges:
            # a correction that made the slowest rank slower (the time inside a range is not
            # uniform: one hub row can carry a fixed cost) is not kept
            self.ranges = best[1]
            self._build_shard()
            self.balance += ', best measured split kept'
        return self.balance_history

    # ---- the recursion -------------------------------------------------------------------
    def run_levels(self, X0: torch.Tensor, levels: int, events: Optional[list] = None):
        """X0: the [n, d_total] level-0 input (every rank holds it).  Returns the last level's
        (sum rows, mean rows) of this rank: rows [lo_r, hi_r), columns [col_lo, col_hi).  With the
        fused exchange the mean rows are a VIEW of this rank's replica, which the peers overwrite
        in their next run_levels: synchronise the ranks before any of them starts another call
        if the views are still in use."""
        d = self.d
        cur = X0[:, self.col_lo:self.col_hi]
        last = None
        # level 0 reads this rank's columns of X0 in place (row stride d_total: no staging copy, the
        # replicas only ever hold levels >= 1) as long as a row still fills whole 128-byte L2
        # lines; narrower column groups gather from a compact copy (measured, C3, d = 16:
        # 8.75 ms in place vs 7.74 ms compact per level)
        if self.C > 1 and self.d * 4 < 128:
            cur = cur.contiguous()
        lo, hi = self.ranges[self.r]
        nvtx = os.environ.get('GR_NVTX') == '1'      # one range per level for nsys / ncu timelines
        for level in range(levels):
            if nvtx:
                if level:
                    torch.cuda.nvtx.range_pop()
                torch.cuda.nvtx.range_push(f'refex level {level + 1} ({self.exchange})')
            if events is not None:
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            if self.R == 1:
                out = self.handle.aggregate(cur, out=self.out[level & 1])
                if events is not None:
                    e1.record()
                    events.append((e0, e1))
                cur = out[:, d:]
                last = (out[:, :d], out[:, d:])
            elif self.exchange == 'peer':
                which = (level + 1) & 1
                self.handle.aggregate_bcast(cur, self.sums, self.peers.replica_ptrs(which), d, lo)
                if events is not None:
                    e1.record()
                self.peers.barrier()
                if events is not None:
                    e2 = torch.cuda.Event(enable_timing=True)
                    e2.record()
                    events.append((e0, e1, e2))       # kernel = e0..e1, barrier wait = e1..e2
                cur = self.peers.replicas[which]
                last = (self.sums, cur[lo:hi])
            else:
                nxt = self.full[level & 1]
                self.handle.aggregate_into(cur, self.sums, nxt[lo:hi])
                if events is not None:
                    e1.record()
                exchange_rows(nxt, self.ranges, self.r, self.group, self.members)
                if events is not None:
                    e2 = torch.cuda.Event(enable_timing=True)
                    e2.record()
                    events.append((e0, e1, e2))
                cur = nxt
                last = (self.sums, nxt[lo:hi])
        if nvtx and levels:
            torch.cuda.nvtx.range_pop()
        if self.peers is not None:
            # a rank that never arrived at a barrier is an error NOW, not at close(): the replicas
            # would be incomplete
            epoch = self.peers.timed_out()
            if epoch:
                raise RuntimeError(f'peer barrier timed out at epoch {epoch}: a rank of the '
                                   f'exchange group did not finish its level')
        return last

    def run_levels_host(self, X0_host: torch.Tensor, levels: int, out_host: torch.Tensor):
        """Host-buffer form (the call a binding inside the reference would make): H2D of this
        rank's rows and columns of X0 from pinned memory, `levels` levels, D2H of the own rows of
        every level -- all inside the library.  out_host: [levels, rows, 2 * d] (sum | mean per
        row) for a whole-graph handle (gr_refex_levels_host_f32), [levels, 2, rows, d] (sum rows,
        then mean rows) for a node-range shard (gr_refex_levels_host_sharded_f32)."""
        if self.R == 1:
            view = X0_host[:, self.col_lo:self.col_hi]
            return self.handle.levels_host(view, levels, 'mean', out_host)
        if self.exchange != 'peer':
            raise RuntimeError('the host-buffer entry point of a node-range shard needs the peer '
                               'exchange (gr_refex_levels_host_sharded_f32)')
        lo, _ = self.ranges[self.r]
        self.peers.epoch = self.handle.levels_host_sharded(
            X0_host, self.col_lo, self.d, levels, lo, self.peers.replica_ptrs(0),
            self.peers.replica_ptrs(1), self.peers.flag_ptrs(), self.peers.index,
            self.peers.epoch, out_host)
        epoch = self.peers.timed_out()
        if epoch:
            raise RuntimeError(f'peer barrier timed out at epoch {epoch}')
        return out_host

    def close(self):
        if self.peers is not None:
            peers, self.peers = self.peers, None
            epoch = peers.timed_out()
            peers.release()                 # unmap and free first, then report
            if epoch:
                raise RuntimeError(f'peer barrier timed out at epoch {epoch}')
