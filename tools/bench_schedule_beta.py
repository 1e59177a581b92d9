#!/usr/bin/env python
"""Schedule beta of SURVEY.md section 8(d) on one GPU: the reference-shaped recursion when the
pruner retains everything -- level l aggregates ALL columns of level l-1's output, so the input
width doubles every level (C2: 32 -> 64 -> 128 -> 256 input columns over 4 levels; total
arc.features = nnz * (32 + 64 + 128 + 256)).  bench.py measures schedule alpha (fixed width);
per-level throughput in arc.features/s and bytes per arc.feature are the same for both, this
script shows it.  Not yet run on hardware in round 1.

    python tools/bench_schedule_beta.py [--workload c2] [--steps 5] [--warmup 3]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='c2', choices=['c2', 'tiny'])
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    args = ap.parse_args()
    device = torch.device('cuda', 0)
    g, d0, levels = bench.build_graph(args.workload, device)
    handle = g.handle(device)
    X0 = torch.rand(g.n, d0, device=device, generator=torch.Generator(device=device).manual_seed(0))
    widths = [d0 * 2 ** l for l in range(levels)]
    outs = [torch.empty((g.n, 2 * w), dtype=torch.float32, device=device) for w in widths]

    def one_step(events=None):
        cur = X0
        for l in range(levels):
            if events is not None:
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            handle.aggregate(cur, out=outs[l])
            if events is not None:
                e1.record()
                events.append((l, e0, e1))
            cur = outs[l]

    for _ in range(max(3, args.warmup)):
        one_step()
    torch.cuda.synchronize()
    events = []
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        one_step(events)
    t1.record()
    torch.cuda.synchronize()
    ms_step = t0.elapsed_time(t1) / args.steps
    peak, src = bench.measured_peak_hbm()
    per_level = []
    for l, w in enumerate(widths):
        ms = sum(a.elapsed_time(b) for k, a, b in events if k == l) / args.steps
        alg = bench.algorithmic_bytes_per_level(g.n, g.nnz, w)
        per_level.append({'input_columns': w, 'ms': ms, 'alg_GBps': alg / ms / 1e6,
                          'frac_of_hbm_peak': alg / ms / 1e6 / peak,
                          'arc_features_per_s': g.nnz * w / ms * 1e3})
    print(json.dumps({
        'metric': bench.METRIC, 'schedule': 'beta: every level aggregates all columns of the '
        'previous level, width doubles', 'workload': bench.WORKLOAD_NAMES[args.workload],
        'value': g.nnz * sum(widths) / ms_step * 1e3, 'unit': bench.UNIT, 'ms_per_step': ms_step,
        'steps': args.steps, 'peak': peak, 'peak_source': src, 'levels': per_level}))


if __name__ == '__main__':
    main()
