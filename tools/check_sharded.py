#!/usr/bin/env python
"""Multi-GPU parity check of the sharded recursion (run under torchrun, one rank per GPU):
    torchrun --nproc-per-node N tools/check_sharded.py [--size 300000] [--depth 4]
(option names are chosen so that torchrun's own abbreviation matching cannot claim them)

For every sharding mode -- node ranges with the fused peer exchange, node ranges with the NCCL
all-gather, column groups x node ranges, measured-time rebalancing, and the host-buffer entry
point -- every rank recomputes the UNSHARDED recursion of its own column group on its own GPU
and compares, bit for bit, the rows it owns (sums and means) and the full replica of every
level's input with what the sharded engine produced; the column-group result is also compared
with the full-width single-GPU recursion at 2e-6 relative (a narrower row changes the number of
lane groups, hence the fp32 summation order)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def unsharded_levels(g, X, levels):
    """Single-GPU recursion on the columns X holds: list of [n, 2d] level outputs."""
    h = g.handle(X.device)
    h.tune_hot_rows(X.shape[1] * 4)
    d = X.shape[1]
    outs, cur = [], X.contiguous()
    for _ in range(levels):
        out = h.aggregate(cur).clone()
        outs.append(out)
        cur = out[:, d:]
    return outs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', dest='n', type=int, default=300_000)
    ap.add_argument('--attach', dest='m', type=int, default=12)
    ap.add_argument('--width', dest='d', type=int, default=64)
    ap.add_argument('--depth', dest='levels', type=int, default=4)
    args = ap.parse_args()
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=device)
    from graphrole_b200 import shard
    from graphrole_b200.graph.generators import barabasi_albert_csr

    g = barabasi_albert_csr(args.n, args.m, seed=5, device=device)
    d = args.d
    X0 = torch.rand(g.n, d, device=device,
                    generator=torch.Generator(device=device).manual_seed(1))
    full_ref = unsharded_levels(g, X0, args.levels)
    failures = 0
    modes = [('peer', 1, False), ('nccl', 1, False), ('peer', 1, True)]
    if world % 2 == 0:
        modes += [('peer', 2, False), ('nccl', 2, False)]
        if world >= 4:
            modes += [('peer', 2, True)]
    if world >= 2:
        modes += [('peer', world, False)]          # column groups only: no exchange at all
    for mode, C, rebalance in modes:
        eng = shard.ShardedRefex(g, d, world=world, rank=rank, group=dist, exchange=mode,
                                 col_groups=C)
        c0, c1 = eng.col_lo, eng.col_hi
        ref_levels = full_ref if C == 1 else unsharded_levels(g, X0[:, c0:c1], args.levels)
        if rebalance:
            eng.autobalance(X0, args.levels, rounds=2, tolerance=1.0)
        lo, hi = eng.ranges[eng.r]
        dl = eng.d
        for levels in range(1, args.levels + 1):      # also exercises replica ping-pong reuse
            sums, means = eng.run_levels(X0, levels)
            torch.cuda.synchronize()
            ref = ref_levels[levels - 1]
            ok_s = torch.equal(sums, ref[lo:hi, :dl])
            ok_m = torch.equal(means, ref[lo:hi, dl:])
            if eng.exchange == 'peer':
                ok_f = torch.equal(eng.peers.replicas[levels & 1], ref[:, dl:])
            elif eng.exchange == 'nccl':
                ok_f = torch.equal(eng.full[(levels - 1) & 1], ref[:, dl:])
            else:
                ok_f = True
            wide = full_ref[levels - 1]
            ok_w = torch.allclose(means, wide[lo:hi, d + c0:d + c1], rtol=2e-6, atol=0) and \
                torch.allclose(sums, wide[lo:hi, c0:c1], rtol=2e-6, atol=0)
            torch.cuda.synchronize()
            dist.barrier()        # the peers overwrite the replicas in their next run_levels
            if not (ok_s and ok_m and ok_f and ok_w):
                failures += 1
                print(f'[rank {rank}] MISMATCH mode={mode}/{eng.exchange} C={C} levels={levels} '
                      f'sum={ok_s} mean={ok_m} replica={ok_f} full-width={ok_w}', flush=True)
        # host-buffer entry point: same bits as the device path, every level
        if eng.exchange in ('peer', 'none'):
            Xh = torch.empty((g.n, d), dtype=torch.float32, pin_memory=True)
            Xh.copy_(X0)
            sharded = eng.R > 1
            outh = torch.empty((args.levels, 2, hi - lo, dl) if sharded else
                               (args.levels, hi - lo, 2 * dl), dtype=torch.float32,
                               pin_memory=True)
            for _ in range(2):                          # twice: buffers and epochs are reused
                outh.fill_(-1.0)
                eng.run_levels_host(Xh, args.levels, outh)
                dist.barrier()
                for level in range(args.levels):
                    ref = ref_levels[level]
                    got = outh[level].to(device)
                    if sharded:
                        got = torch.cat([got[0], got[1]], dim=1)
                    if not torch.equal(got, ref[lo:hi]):
                        failures += 1
                        print(f'[rank {rank}] MISMATCH host path mode={mode} C={C} '
                              f'level={level}', flush=True)
        note = f' ({eng.exchange_note})' if eng.exchange_note else ''
        print(f'[rank {rank}] mode={mode} C={C} rebalance={rebalance} -> ran as {eng.exchange}'
              f'{note}: rows [{lo}, {hi}) cols [{c0}, {c1}) nnz={eng.local_nnz} '
              f'{eng.balance_history[-1:] if rebalance else ""}', flush=True)
        eng.close()
    t = torch.tensor([failures], device=device)
    dist.all_reduce(t)
    if rank == 0:
        print('SHARDED CHECK', 'OK' if int(t.item()) == 0 else f'FAILED ({int(t.item())})',
              flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 0 else 1)


if __name__ == '__main__':
    main()
