#!/usr/bin/env python
"""Multi-GPU parity check of the node-range sharded recursion (run under torchrun, one rank per
GPU):  torchrun --nproc-per-node N tools/check_sharded.py [--size 300000] [--depth 4]
(option names are chosen so that torchrun's own abbreviation matching cannot claim them)

Every rank computes the unsharded recursion on its own GPU and compares, bit for bit, the rows
it owns (sums and means) and the full replica of every level's input with what the sharded
engine produced -- for both exchange forms ('peer': fused gather + NVLink stores + flag
barrier; 'nccl': all-gather)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


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
    X0 = torch.rand(g.n, args.d, device=device,
                    generator=torch.Generator(device=device).manual_seed(1))
    single = shard.ShardedRefex(g, args.d)
    ref_levels = []
    cur = X0
    for _ in range(args.levels):
        out = single.handle.aggregate(cur).clone()
        ref_levels.append(out)
        cur = out[:, args.d:]
    failures = 0
    for mode in ('peer', 'nccl'):
        eng = shard.ShardedRefex(g, args.d, world=world, rank=rank, group=dist, exchange=mode)
        lo, hi = eng.ranges[rank]
        for levels in range(1, args.levels + 1):      # also exercises replica ping-pong reuse
            sums, means = eng.run_levels(X0, levels)
            torch.cuda.synchronize()
            ref = ref_levels[levels - 1]
            ok_s = torch.equal(sums, ref[lo:hi, :args.d])
            ok_m = torch.equal(means, ref[lo:hi, args.d:])
            full = (eng.peers.replicas[levels & 1] if eng.exchange == 'peer'
                    else eng.full[(levels - 1) & 1])
            ok_f = torch.equal(full, ref[:, args.d:])
            if not (ok_s and ok_m and ok_f):
                failures += 1
                print(f'[rank {rank}] MISMATCH mode={mode}/{eng.exchange} levels={levels} '
                      f'sum={ok_s} mean={ok_m} replica={ok_f}', flush=True)
        note = f' ({eng.exchange_note})' if eng.exchange_note else ''
        print(f'[rank {rank}] mode={mode} -> ran as {eng.exchange}{note}: rows [{lo}, {hi}) '
              f'nnz={eng.local_nnz}', flush=True)
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
