#!/usr/bin/env python
"""Tuning sweep for the ReFeX gather kernel (development tool, GPU box only): times one level on
a benchmark workload for combinations of the kernel's tuning knobs (read from the environment
by gr_refex_aggregate_f32 on every call)."""
import argparse
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='c3')
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--unroll', default='2,4,8')
    ap.add_argument('--rows', default='4,8,16')
    ap.add_argument('--extra', default='', help='NAME=v1,v2 additional env knob to sweep')
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    g, d, levels = bench.build_graph(args.workload, dev)
    X = torch.rand(g.n, d, device=dev)
    out = torch.empty(g.n, 2 * d, device=dev)
    h = g.handle(dev)
    alg = bench.algorithmic_bytes_per_level(g.n, g.nnz, d)
    extra_name, extra_vals = None, ['']
    if args.extra:
        extra_name, vals = args.extra.split('=')
        extra_vals = vals.split(',')
    results = []
    for u, r, ev in itertools.product(args.unroll.split(','), args.rows.split(','), extra_vals):
        os.environ['GR_REFEX_UNROLL'] = u
        os.environ['GR_REFEX_ROWS_PER_WARP'] = r
        if extra_name:
            os.environ[extra_name] = ev
        for _ in range(2):
            h.aggregate(X, out=out)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            h.aggregate(X, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        rec = {'unroll': int(u), 'rows_per_warp': int(r), 'ms': ms, 'GBps': alg / ms / 1e6}
        if extra_name:
            rec[extra_name] = ev
        results.append(rec)
        print(json.dumps(rec), flush=True)
    best = min(results, key=lambda x: x['ms'])
    print('BEST', json.dumps(best))


if __name__ == '__main__':
    main()
