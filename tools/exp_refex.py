#!/usr/bin/env python
"""Driver of tools/exp_refex.cu (development harness, GPU box only)."""
import argparse
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

SO = os.path.join(ROOT, 'tools', 'libexp_refex.so')


def build():
    src = os.path.join(ROOT, 'tools', 'exp_refex.cu')
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.run(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
                        '-std=c++17', '-Xcompiler', '-fPIC', '-shared', '-o', SO, src], check=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='c3')
    ap.add_argument('--variants', default='0,1,2')
    ap.add_argument('--rows', default='16')
    ap.add_argument('--hot', default='0')
    ap.add_argument('--reps', type=int, default=5)
    args = ap.parse_args()
    build()
    lib = ctypes.CDLL(SO)
    lib.exp_run.restype = ctypes.c_float
    lib.exp_run.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                            ctypes.c_int, ctypes.c_int, ctypes.c_int]
    dev = torch.device('cuda', 0)
    g, d, levels = bench.build_graph(args.workload, dev)
    X = torch.rand(g.n, d, device=dev)
    ref = g.handle(dev).aggregate(X)
    deg = g.out_degree()
    ok_rows = deg <= 2048
    out = torch.zeros(g.n, 2 * d, device=dev)
    alg = bench.algorithmic_bytes_per_level(g.n, g.nnz, d)
    for v in [int(x) for x in args.variants.split(',')]:
        for rows in [int(x) for x in args.rows.split(',')]:
            for hot in [int(x) for x in args.hot.split(',')]:
                if hot and not (30 <= v < 40):
                    continue
                out.zero_()
                ms = lib.exp_run(v, d, g.rowptr.data_ptr(), g.colidx.data_ptr(), X.data_ptr(), d,
                                 g.n, out.data_ptr(), rows, hot, args.reps)
                torch.cuda.synchronize()
                sel = ok_rows if v < 100 else torch.ones_like(ok_rows)
                err = float(((out[sel] - ref[sel]).abs() / (ref[sel].abs() + 1e-6)).max())
                print(json.dumps({'variant': v, 'rows_per_warp': rows, 'hot': hot,
                                  'ms': round(ms, 4), 'GBps': round(alg / ms / 1e6, 1),
                                  'max_rel_err': err}), flush=True)


if __name__ == '__main__':
    main()
