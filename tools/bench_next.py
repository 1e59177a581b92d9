#!/usr/bin/env python
"""Timing of the "next" rows at benchmark scale on one B200:
level-0 features of the C3 graph, vertical log binning and pairwise gaps of a 10 M-row feature
matrix.  CUDA events after one warm-up call; prints one JSON line per measurement."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from graphrole_b200 import _native
from graphrole_b200.graph import level0
from graphrole_b200.graph.generators import barabasi_albert_csr


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=10_000_000)
    ap.add_argument('--m', type=int, default=20)
    ap.add_argument('--d', type=int, default=64)
    ap.add_argument('--pairs-d', type=int, default=128)
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    g = barabasi_albert_csr(args.n, args.m, seed=0, device=dev)
    ms = timed(lambda: level0.device_features(g), reps=1)
    print(json.dumps({'what': 'level0 (degree, internal_edges, external_edges)', 'n': g.n,
                      'nnz': g.nnz, 'max_degree': int(g.out_degree().max()), 'ms': round(ms, 2),
                      'arcs_per_s': round(g.nnz / ms * 1e3)}), flush=True)
    handle = g.handle(dev)
    X0 = torch.rand(g.n, args.d, device=dev)
    feats = handle.aggregate(X0)                      # a realistic [n, 2d] level output
    del X0
    p = _native.Pruner(g.n, dev)
    bins = torch.empty((feats.shape[1], g.n), dtype=torch.int32, device=dev)
    ms = timed(lambda: p.bin_columns(feats, out=bins), reps=1)
    nbytes = feats.numel() * 4
    print(json.dumps({'what': 'vertical_log_binning fp32', 'n': g.n, 'columns': feats.shape[1],
                      'ms': round(ms, 2), 'ms_per_column': round(ms / feats.shape[1], 3),
                      'keys_per_s': round(feats.numel() / ms * 1e3),
                      'input_GBps': round(nbytes / ms / 1e6, 1)}), flush=True)
    d = min(args.pairs_d, bins.shape[0])
    ms = timed(lambda: p.pairwise_gaps(bins[:d]), reps=2)
    pair_rows = d * (d - 1) / 2 * g.n
    print(json.dumps({'what': 'pairwise Chebyshev gaps', 'n': g.n, 'columns': d,
                      'ms': round(ms, 2), 'pair_rows_per_s': round(pair_rows / ms * 1e3),
                      'bins_GBps_one_read': round(d * g.n * 4 / ms / 1e6, 1)}), flush=True)
    p.close()


if __name__ == '__main__':
    main()
