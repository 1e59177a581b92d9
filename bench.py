#!/usr/bin/env python
"""Benchmark of the ReFeX neighbourhood aggregation (hot path A) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|tiny]
    python bench.py --impl reference ...          # CPU arm (port of the reference's pandas chain)
    torchrun --nproc-per-node N bench.py --gpus N ...   # node-range sharded, one rank per GPU

Metric (BASELINE.json): aggregated edges*features per second =  nnz * d * levels / t, with nnz the
CSR arcs traversed (2|E| for an undirected graph), d the columns aggregated per level.
A step = `levels` recursion levels over the whole graph (schedule "alpha": every level
aggregates d input columns into d sums + d means and the next level recurses on the mean block).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (family, n, m-parameter, d, levels)     -- BASELINE.json configs[2] / configs[1]
    'c3': ('ba', 10_000_000, 20, 64, 5),
    'c2': ('er', 1_000_000, 20_000_000, 32, 4),
    'tiny': ('ba', 200_000, 20, 64, 5),
}
WORKLOAD_NAMES = {
    'c3': 'power-law (Barabasi-Albert) |V|=10M |E|~200M (CSR nnz~400M), 64 features, 5 levels',
    'c2': 'Erdos-Renyi |V|=1M |E|~20M (CSR nnz~40M), 32 features, 4 levels',
    'tiny': 'Barabasi-Albert |V|=200k m=20, 64 features, 5 levels (development only)',
}
# dram__bytes_read.sum + dram__bytes_write.sum of one refex_gather_kernel launch, from the
# `ncu --set full` capture committed as profiles/r1_ncu_refex_final_raw.csv (single GPU only)
NCU_TRAFFIC_BYTES = {'c3': 91.29e9 + 5.49e9}
METRIC = 'refex_aggregated_edges_x_features_per_sec'
UNIT = 'arc*features/s'


def build_graph(workload, device):
    from graphrole_b200.graph.generators import barabasi_albert_csr, erdos_renyi_csr
    family, n, m, d, levels = WORKLOADS[workload]
    if family == 'ba':
        g = barabasi_albert_csr(n, m, seed=0, device=device)
    else:
        g = erdos_renyi_csr(n, m, seed=0, device=device)
    return g, d, levels


def algorithmic_bytes_per_level(n_rows, n_nnz, d):
    """gathered rows + colidx + rowptr + sum/mean writes (fp32)."""
    return n_nnz * d * 4 + n_nnz * 4 + (n_rows + 1) * 8 + n_rows * 2 * d * 4


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                 '-lms', '100', '-i', str(self.gpu_index)],
                stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if not self.path or not os.path.exists(self.path):
            return out
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in open(self.path):
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[4:8]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(smax),
                       reasons=sorted(reasons), samples=len(sm))
        return out


def measured_peak_hbm():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except (OSError, KeyError, ValueError):
        return 6650.0, 'fallback (B200_PROFILING.md)'


# ---- CPU baselines (oracle/ is test + baseline infrastructure; never on the product path) ----

def cpu_pandas_port(rp, ci, X_host, d, seconds, rng):
    """The reference's per-node pandas chain (extract.py:105-118) on sampled nodes."""
    import pandas as pd
    from oracle import refex_oracle
    n = rp.shape[0] - 1
    feats = pd.DataFrame(X_host.astype(np.float64), columns=[f'f{j}' for j in range(d)])
    cols = list(feats.columns)
    probe = rng.choice(n, 4, replace=False)
    t0 = time.perf_counter()
    refex_oracle.pandas_chain_rows(feats, cols, probe, rp, ci)
    per_node = (time.perf_counter() - t0) / len(probe)
    k = int(max(8, min(n, seconds / max(per_node, 1e-6))))
    rows = rng.choice(n, k, replace=False)
    t0 = time.perf_counter()
    refex_oracle.pandas_chain_rows(feats, cols, rows, rp, ci)
    dt = time.perf_counter() - t0
    arcs = int((rp[rows + 1] - rp[rows]).sum())
    return arcs * d / dt, k, arcs, dt


def cpu_fair_comparator(rp, ci, X_host, d, rows):
    """Multi-threaded C restatement (fp64 accumulation) on a contiguous block of rows."""
    from oracle import refex_oracle
    sel = np.arange(rows, dtype=np.int64)
    refex_oracle.aggregate_rows_c(sel[:1000], rp, ci, X_host)   # warm
    t0 = time.perf_counter()
    refex_oracle.aggregate_rows_c(sel, rp, ci, X_host)
    dt = time.perf_counter() - t0
    arcs = int(rp[rows] - rp[0])
    return arcs * d / dt, refex_oracle.c_threads(), arcs, dt


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU implementation of the path.  The reference is
    pure Python and /root/reference does not exist on the GPU box, so this times the oracle's
    faithful port of its per-node pandas chain (single Python thread, like the reference)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    device = 'cuda' if torch.cuda.is_available() else 'cpu'
    g, d, levels = build_graph(args.workload, device)
    rp, ci = g.host_arrays()
    n = g.n
    X_host = torch.rand(n, d, generator=torch.Generator().manual_seed(0)).numpy()
    rng = np.random.RandomState(0)
    per_step_seconds = max(1.0, min(6.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_pandas_port(rp, ci, X_host, d, per_step_seconds, rng)
    rates, nodes, times = [], 0, []
    for _ in range(args.steps):
        rate, k, arcs, dt = cpu_pandas_port(rp, ci, X_host, d, per_step_seconds, rng)
        rates.append(rate)
        nodes += k
        times.append(dt)
    value = float(np.mean(rates))
    fair, threads, _, _ = cpu_fair_comparator(rp, ci, X_host, d, min(n, 1_000_000))
    sample = (f'{nodes // max(1, args.steps)} uniformly sampled nodes per step of the same graph '
              f'and feature matrix, one recursion level, rate not extrapolated')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * float(np.mean(times)), 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD_NAMES[args.workload], 'n': n, 'nnz': g.nnz, 'd': d,
                   'levels': levels},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                         'sample': sample, 'host_cores': os.cpu_count(),
                         'fair_c_openmp_f64': {'value': fair, 'unit': UNIT, 'cores': threads}},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ---- GPU arm ----------------------------------------------------------------------------------

def run_gpu_arm(args):
    from graphrole_b200 import _native
    from graphrole_b200 import shard as shard_mod

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('bench.py --gpus N>1 must be launched with torchrun '
                             '(one rank per GPU)')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=device)
    _native.load()

    g, d, levels = build_graph(args.workload, device)
    n, nnz = g.n, g.nnz
    X0 = torch.rand(n, d, device=device, generator=torch.Generator(device=device).manual_seed(0))

    engine = shard_mod.ShardedRefex(g, d, world=world, rank=rank, group=dist)
    stream = torch.cuda.current_stream()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    level_events = []

    def one_step(record):
        engine.run_levels(X0, levels, level_events if record else None)

    for _ in range(args.warmup):
        one_step(False)
    barrier()
    launches0 = _native.launch_count()
    with ClockSampler(local_rank) as clocks:
        barrier()
        t_start = torch.cuda.Event(enable_timing=True)
        t_stop = torch.cuda.Event(enable_timing=True)
        t_start.record(stream)
        for _ in range(args.steps):
            one_step(True)
        t_stop.record(stream)
        barrier()
    launches = _native.launch_count() - launches0
    ms_total = t_start.elapsed_time(t_stop)
    if dist is not None:
        t = torch.tensor([ms_total], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = nnz * d * levels / (ms_per_step * 1e-3)

    # dominant kernel: the gather-reduce launch of each level (events bracket one
    # gr_refex_aggregate_f32 call = gather kernel + the microsecond-scale hub fix-up)
    kern_ms = [a.elapsed_time(b) for a, b in level_events]
    kern_ms_avg = float(np.mean(kern_ms)) if kern_ms else float('nan')
    peak, peak_src = measured_peak_hbm()
    alg_bytes = algorithmic_bytes_per_level(engine.local_rows, engine.local_nnz, d)
    nvlink_bytes = 0
    if engine.exchange == 'peer':        # the mean rows also go to the world-1 peer replicas
        nvlink_bytes = engine.local_rows * d * 4 * (world - 1)
    achieved = alg_bytes / (kern_ms_avg * 1e-3) / 1e9

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD_NAMES[args.workload], 'n': n, 'nnz': nnz,
                   'undirected_edges': nnz // 2, 'd': d, 'levels': levels,
                   'schedule': 'alpha: fixed width, recurse on the mean block, pruning disabled',
                   'l2': 'inputs larger than L2 (feature matrix %.2f GB vs 126 MB L2)'
                         % (n * d * 4 / 1e9),
                   'parallelism': 'single GPU' if world == 1 else
                   f'node-range sharded x{world}, {engine.balance}, full input replica per GPU',
                   'exchange': {'none': 'none (single GPU)',
                                'peer': 'fused: gather kernel stores mean rows into every '
                                        'replica over NVLink-mapped peer pointers + flag barrier',
                                'nccl': 'all-gather of the mean rows after the kernel'}
                   [engine.exchange] + (f' [{engine.exchange_note}]' if engine.exchange_note
                                        else ''),
                   'value_with_undirected_edge_convention': value / 2},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                     'frac': achieved / peak,
                     'traffic': NCU_TRAFFIC_BYTES.get(args.workload) if world == 1 else None,
                     'peak_source': peak_src,
                     'kernel': 'refex_gather_kernel' if engine.exchange != 'peer'
                     else 'refex_gather_bcast_kernel', 'kernel_ms_avg': kern_ms_avg,
                     'algorithmic_bytes_per_launch': alg_bytes,
                     'nvlink_store_bytes_per_launch': nvlink_bytes},
        'clocks': clocks.summary(),
        'gpu_launches': launches,
    }

    engine.close()
    # ---- e2e: host buffers through the C-ABI, copies inside the timed region (N = 1) -------
    if world == 1 and not args.no_e2e:
        del engine
        torch.cuda.empty_cache()
        h = g.handle(device)
        Xh = torch.empty((n, d), dtype=torch.float32, pin_memory=True)
        Xh.copy_(X0)
        outh = torch.empty((levels, n, 2 * d), dtype=torch.float32, pin_memory=True)
        h.levels_host(Xh, levels, 'mean', outh)       # warm-up (allocates staging)
        e2e_steps = max(1, min(args.steps, 3))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            h.levels_host(Xh, levels, 'mean', outh)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e2e_steps
        line['e2e'] = {'value': nnz * d * levels / dt, 'unit': UNIT,
                       'h2d_bytes_per_step': n * d * 4,
                       'd2h_bytes_per_step': levels * n * 2 * d * 4,
                       'ms_per_step': dt * 1e3, 'steps': e2e_steps,
                       'api': 'gr_refex_levels_host_f32 (pinned host X in, all levels out)'}
        del outh

    # ---- CPU baseline beside it (rank 0, N = 1 only) -----------------------------------------
    if world == 1 and not args.no_cpu_baseline:
        rp, ci = g.host_arrays()
        X_host = X0.cpu().numpy()
        rng = np.random.RandomState(0)
        rate, k, arcs, dt = cpu_pandas_port(rp, ci, X_host, d, 12.0, rng)
        fair, threads, farcs, fdt = cpu_fair_comparator(rp, ci, X_host, d, min(n, 1_000_000))
        line['cpu_baseline'] = {
            'value': rate, 'unit': UNIT, 'cores': 1, 'kind': 'port',
            'sample': f'{k} uniformly sampled nodes ({arcs} arcs) of the same graph/features '
                      f'through the reference\'s per-node pandas chain, {dt:.1f} s',
            'host_cores': os.cpu_count(),
            'fair_c_openmp_f64': {'value': fair, 'unit': UNIT, 'cores': threads,
                                  'sample': f'first {min(n, 1_000_000)} rows ({farcs} arcs), '
                                            f'{fdt:.2f} s'}}

    # ---- the "next" rows of the path on the same graph (N = 1) ----------
    if world == 1 and not args.no_next:
        try:
            line['next'] = next_rows_section(g, d, device)
        except Exception as exc:
            line['next'] = {'error': repr(exc)}

    # ---- hot path B beside it (N = 1): RolX NMF, BASELINE.json configs[4] ---------------------
    if world == 1 and not args.no_nmf:
        try:
            del X0
            torch.cuda.empty_cache()
            line['nmf'] = nmf_section(device)
        except Exception as exc:      # never lose the headline line to the secondary section
            line['nmf'] = {'error': repr(exc)}

    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def next_rows_section(g, d, device):
    """What runs around the aggregation in a whole extract_features(): level-0 features from the
    CSR, vertical log binning of one level's [n, 2d] output, pairwise Chebyshev gaps of the binned
    columns, and the whole device-resident recursion (pruning enabled, so the widths are the
    data's own).  CUDA events, one warm-up call each."""
    from graphrole_b200 import _native
    from graphrole_b200.features.device import DeviceRecursiveFeatureExtractor
    from graphrole_b200.graph import level0

    def timed(fn, reps=1):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    out = {}
    ms = timed(lambda: level0.device_features(g))
    out['level0'] = {'ms': ms, 'arcs_per_s': g.nnz / ms * 1e3,
                     'columns': ['degree', 'internal_edges', 'external_edges']}
    feats = g.handle(device).aggregate(torch.rand(g.n, d, device=device))
    pruner = _native.Pruner(g.n, device)
    bins = torch.empty((feats.shape[1], g.n), dtype=torch.int32, device=device)
    ms = timed(lambda: pruner.bin_columns(feats, out=bins))
    out['vertical_log_binning'] = {'columns': feats.shape[1], 'ms': ms,
                                   'keys_per_s': feats.numel() / ms * 1e3}
    ms = timed(lambda: pruner.pairwise_gaps(bins), reps=2)
    f = bins.shape[0]
    out['pairwise_gaps'] = {'columns': f, 'ms': ms,
                            'pair_rows_per_s': f * (f - 1) / 2 * g.n / ms * 1e3}
    pruner.close()
    del feats, bins
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    rfe = DeviceRecursiveFeatureExtractor(g, device=device)
    names, values = rfe.extract_features_device()
    torch.cuda.synchronize()
    out['extract_features_device_resident'] = {
        'wall_s': time.perf_counter() - t0, 'generations': rfe.generation_count,
        'features': len(names), 'kernel_ms': {k: round(v, 2) for k, v in rfe.timings_ms.items()}}
    return out


def nmf_section(device, n=10_000_000, f=512, ranks=(4, 8, 16, 32), iters=10):
    """ms per multiplicative-update iteration on X = 10M x 512 fp32 (synthetic U[0,1)), tcgen05
    path, tol = 0 (no convergence pass in the timed region); algorithmic bytes n*f*4 + 2*n*r*4."""
    from graphrole_b200.roles import factor
    gen = torch.Generator(device=device).manual_seed(0)
    X = torch.rand(n, f, device=device, generator=gen)
    peak, _ = measured_peak_hbm()
    rows = []
    for r in ranks:
        W = torch.rand(n, r, device=device, generator=gen) + 0.1
        H = torch.rand(r, f, device=device, generator=gen) + 0.1
        solver = factor.NmfSolver(n, f, r, device)
        solver.update(X, W, H, max_iter=3, tol=0, want_error=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        solver.update(X, W, H, max_iter=iters, tol=0, want_error=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        alg = n * f * 4 + 2 * n * r * 4
        flops = 4 * n * f * r + 4 * n * r * r + 2 * r * r * f
        rows.append({'r': r, 'path': solver.last_path, 'ms_per_iter': ms,
                     'alg_GBps': alg / ms / 1e6, 'frac_of_hbm_peak': alg / ms / 1e6 / peak,
                     'alg_TFLOPs': flops / ms / 1e9})
        solver.close()
        del W, H
    out = {'workload': f'X {n}x{f} fp32 U[0,1), shared random init, {iters} iterations',
           'dtype': 'tf32 MMA / fp32 accumulate', 'per_rank': rows}
    try:      # sklearn MU (the reference's solver) on a row sample, all BLAS threads
        from sklearn.decomposition import _nmf as sk
        m = 200_000
        Xc = X[:m].double().cpu().numpy()
        rng = np.random.RandomState(0)
        W0, H0 = rng.rand(m, 32) + 0.1, rng.rand(32, f) + 0.1
        t0 = time.perf_counter()
        sk._fit_multiplicative_update(Xc, W0, H0, 'frobenius', max_iter=2, tol=0)
        dt = (time.perf_counter() - t0) / 2
        out['cpu_sklearn_f64'] = {'r': 32, 'rows': m, 's_per_iter_sample': dt,
                                  's_per_iter_scaled_to_n': dt * n / m,
                                  'cores': os.cpu_count()}
    except Exception as exc:
        out['cpu_sklearn_f64'] = {'error': repr(exc)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c3', choices=list(WORKLOADS))
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-nmf', action='store_true')
    ap.add_argument('--no-next', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_gpu_arm(args)


if __name__ == '__main__':
    main()
