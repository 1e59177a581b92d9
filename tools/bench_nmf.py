#!/usr/bin/env python
"""NMF (hot path B) benchmark: BASELINE.json configs[4] -- X = 10M x 512 fp32, r in {4,8,16,32}.

Prints one JSON line per rank r: ms/iteration (CUDA events over a fixed number of iterations,
tol = 0 so no convergence pass is inside the timed region), algorithmic bytes
n*f*4 + 2*n*r*4 per iteration against the measured HBM peak, and algorithmic TFLOP/s
(4nfr + 4nr^2 + 2r^2f per iteration).  --cpu adds sklearn's MU solver on a row sample.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from graphrole_b200 import _native
from graphrole_b200.roles import factor

# development only: time a kernel variant built by tools/build_variant.sh
if os.environ.get('GR_EXP_LIB'):
    _native.LIB_PATH = os.path.abspath(os.environ['GR_EXP_LIB'])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=10_000_000)
    ap.add_argument('--f', type=int, default=512)
    ap.add_argument('--ranks', default='4,8,16,32')
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--paths', default='tcgen05,ffma')
    ap.add_argument('--cpu', action='store_true')
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    gen = torch.Generator(device=dev).manual_seed(0)
    X = torch.rand(args.n, args.f, device=dev, generator=gen)
    peak, src = bench.measured_peak_hbm()
    for r in [int(v) for v in args.ranks.split(',')]:
        W0 = torch.rand(args.n, r, device=dev, generator=gen) + 0.1
        H0 = torch.rand(r, args.f, device=dev, generator=gen) + 0.1
        for path in args.paths.split(','):
            use_tf32 = path == 'tcgen05'
            iters = args.iters if use_tf32 else max(2, args.iters // 5)
            solver = factor.NmfSolver(args.n, args.f, r, dev)
            W, H = W0.clone(), H0.clone()
            solver.update(X, W, H, max_iter=3, tol=0, use_tf32=use_tf32, want_error=False)
            torch.cuda.synchronize()
            l0 = _native.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            solver.update(X, W, H, max_iter=iters, tol=0, use_tf32=use_tf32, want_error=False)
            e1.record()
            torch.cuda.synchronize()
            launches = _native.launch_count() - l0
            ms = e0.elapsed_time(e1) / iters
            _, err = solver.update(X, W, H, max_iter=0, tol=0, use_tf32=use_tf32, want_error=True)
            # one convergence check alone (max_iter = 0, tol > 0: only the error at init runs),
            # and sklearn's schedule: 20 iterations with a check every 10
            solver.update(X, W, H, max_iter=0, tol=1e-30, use_tf32=use_tf32)
            e0.record()
            for _ in range(3):
                solver.update(X, W, H, max_iter=0, tol=1e-30, use_tf32=use_tf32)
            e1.record()
            torch.cuda.synchronize()
            ms_check = e0.elapsed_time(e1) / 3
            e0.record()
            n_it, err_c = solver.update(X, W, H, max_iter=20, tol=1e-30, check_every=10,
                                        use_tf32=use_tf32)
            e1.record()
            torch.cuda.synchronize()
            ms_checked = e0.elapsed_time(e1) / max(n_it, 1)
            alg_bytes = args.n * args.f * 4 + 2 * args.n * r * 4
            flops = 4 * args.n * args.f * r + 4 * args.n * r * r + 2 * r * r * args.f
            print(json.dumps({
                'path': solver.last_path, 'n': args.n, 'f': args.f, 'r': r, 'iters': iters,
                'ms_per_iter': round(ms, 3), 'alg_GBps': round(alg_bytes / ms / 1e6, 1),
                'frac_of_hbm_peak': round(alg_bytes / ms / 1e6 / peak, 3), 'peak': peak,
                'alg_TFLOPs': round(flops / ms / 1e9, 2), 'error': err,
                'launches_per_iter': launches / iters, 'ms_per_check': round(ms_check, 3),
                'ms_per_iter_with_checks': round(ms_checked, 3), 'iters_with_checks': n_it,
                'error_at_check': err_c,
                'error_pass': 'ffma' if os.environ.get('GR_NMF_ERROR_FFMA') or
                solver.last_path != 'tcgen05' else 'tcgen05'}), flush=True)
            solver.close()
            del W, H
        del W0, H0
    if args.cpu:
        from sklearn.decomposition import _nmf as sk
        rows = min(args.n, 1_000_000)
        Xc = X[:rows].cpu().numpy().astype(np.float64)
        for r in [int(v) for v in args.ranks.split(',')]:
            rng = np.random.RandomState(0)
            W0, H0 = rng.rand(rows, r) + 0.1, rng.rand(r, args.f) + 0.1
            t0 = time.perf_counter()
            sk._fit_multiplicative_update(Xc, W0, H0, 'frobenius', max_iter=3, tol=0)
            dt = (time.perf_counter() - t0) / 3
            print(json.dumps({'path': 'sklearn_cpu_f64', 'rows': rows, 'f': args.f, 'r': r,
                              's_per_iter_sample': round(dt, 3),
                              's_per_iter_scaled_to_n': round(dt * args.n / rows, 2),
                              'cores': os.cpu_count()}), flush=True)


if __name__ == '__main__':
    main()
