#!/usr/bin/env python
"""Benchmark of the ReFeX neighbourhood aggregation (hot path A) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|tiny]
    python bench.py --impl reference ...          # CPU arm: the reference's own _get_next_features
    torchrun --nproc-per-node N bench.py --gpus N ...   # sharded, one rank per GPU

Metric (BASELINE.json): aggregated edges*features per second =  nnz * d * levels / t, with nnz the
CSR arcs traversed (2|E| for an undirected graph), d the columns aggregated per level.
A step = `levels` recursion levels over the whole graph (schedule "alpha": every level
aggregates d input columns into d sums + d means and the next level recurses on the mean block).
Prints ONE JSON line on rank 0.  What was timed is then VERIFIED outside the timed region
(`parity`): against the float64 C oracle on sampled rows of every level and, sharded, bit for bit
against the unsharded kernel on the same GPU.
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
# `ncu --set full` capture committed under profiles/ (single GPU only; not re-measured per run)
NCU_TRAFFIC = {'c3': {'bytes': 91.148e9 + 5.493e9, 'source': 'profiles/r2_ncu_refex_gather_d64_raw.csv'}}
METRIC = 'refex_aggregated_edges_x_features_per_sec'
UNIT = 'arc*features/s'
RTOL = 1e-5        # north_star: feature matrices within 1e-5 relative (fp32 vs the float64 path)


def build_graph(workload, device):
    from graphrole_b200.graph.generators import barabasi_albert_csr, erdos_renyi_csr
    family, n, m, d, levels = WORKLOADS[workload]
    if family == 'ba':
        g = barabasi_albert_csr(n, m, seed=0, device=device)
    else:
        g = erdos_renyi_csr(n, m, seed=0, device=device)
    return g, d, levels


def workload_config(workload, n, nnz, d, levels):
    """The `config` object: only what defines the workload, so that both arms print the same."""
    return {'workload': WORKLOAD_NAMES[workload], 'n': n, 'nnz': nnz, 'undirected_edges': nnz // 2,
            'd': d, 'levels': levels,
            'schedule': 'alpha: fixed width, recurse on the mean block, pruning disabled',
            'l2': 'inputs larger than L2 (feature matrix %.2f GB vs 126 MB L2)' % (n * d * 4 / 1e9)}


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

class CpuReference:
    """The reference's CPU implementation of the path on sampled nodes of the same graph and
    feature matrix: the UNMODIFIED reference's `_get_next_features` when oracle/_ref holds the
    staged package (kind 'reference'), else the oracle's port of its per-node pandas chain
    (kind 'port').  One Python thread either way -- the reference is a per-node Python loop."""

    def __init__(self, rp, ci, X_host, d):
        from oracle import ref_arm
        self.rp, self.ci, self.d = rp, ci, d
        self.n = rp.shape[0] - 1
        self.level = None
        self.kind = 'port'
        if ref_arm.available():
            try:
                self.level = ref_arm.ReferenceLevel(rp, ci, X_host)
                self.kind = 'reference'
            except Exception as exc:                  # staged copy unusable: say so, use the port
                self.note = f'oracle/_ref unusable ({exc!r}); timing the port'
        if self.level is None:
            import pandas as pd
            self.feats = pd.DataFrame(X_host.astype(np.float64),
                                      columns=[f'f{j}' for j in range(d)])

    def run(self, rows):
        """(rate, arcs, seconds) for one pass over `rows`."""
        if self.level is not None:
            return self.level.timed(rows)
        from oracle import refex_oracle
        t0 = time.perf_counter()
        refex_oracle.pandas_chain_rows(self.feats, list(self.feats.columns), rows, self.rp, self.ci)
        dt = time.perf_counter() - t0
        arcs = int((self.rp[rows + 1] - self.rp[rows]).sum())
        return arcs * self.d / dt, arcs, dt

    def sample(self, seconds, rng):
        """One bounded sample: as many uniformly drawn nodes as fit `seconds` of CPU time."""
        probe = rng.choice(self.n, 4, replace=False)
        self.run(probe)                                # builds pandas' index engine once
        _, _, dt = self.run(probe)
        k = int(max(8, min(self.n, seconds / max(dt / len(probe), 1e-6))))
        rows = rng.choice(self.n, k, replace=False)
        rate, arcs, dt = self.run(rows)
        return rate, k, arcs, dt

    def describe(self):
        return ('graphrole.RecursiveFeatureExtractor._get_next_features of the unmodified '
                'reference (oracle/_ref)' if self.kind == 'reference' else
                'port of the reference\'s per-node pandas chain (oracle.refex_oracle)')


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
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host
    cores, same workload / metric / unit, each step a bounded sample of nodes."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    device = 'cuda' if torch.cuda.is_available() else 'cpu'
    g, d, levels = build_graph(args.workload, device)
    rp, ci = g.host_arrays()
    n = g.n
    X_host = torch.rand(n, d, generator=torch.Generator().manual_seed(0)).numpy()
    rng = np.random.RandomState(0)
    ref = CpuReference(rp, ci, X_host, d)
    per_step_seconds = max(1.0, min(6.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        ref.sample(per_step_seconds, rng)
    rates, nodes, times = [], 0, []
    for _ in range(args.steps):
        rate, k, arcs, dt = ref.sample(per_step_seconds, rng)
        rates.append(rate)
        nodes += k
        times.append(dt)
    value = float(np.mean(rates))
    fair, threads, _, _ = cpu_fair_comparator(rp, ci, X_host, d, min(n, 1_000_000))
    sample = (f'{nodes // max(1, args.steps)} uniformly sampled nodes per step of the same graph '
              f'and feature matrix through {ref.describe()}, one recursion level, rate not '
              f'extrapolated')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * float(np.mean(times)), 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.workload, n, g.nnz, d, levels),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': 1, 'kind': ref.kind,
                         'sample': sample, 'host_cores': os.cpu_count(),
                         'fair_c_openmp_f64': {'value': fair, 'unit': UNIT, 'cores': threads}},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ---- GPU arm ----------------------------------------------------------------------------------

def verify_parity(engine, g, X0, levels, dist, device, rows_per_level=20_000):
    """What was timed, checked outside the timed region.

    * every level: `rows_per_level` sampled rows of this rank's range plus its hub rows against
      the float64 C oracle fed with the SAME level input the GPU used (max relative error, must
      be <= 1e-5), and mean == correctly rounded sum / degree on every row;
    * sharded: this rank re-runs the UNSHARDED kernel on its column group on its own GPU and
      its rows of every level (and, fused exchange, its whole replica) must be bit-identical."""
    from oracle import refex_oracle
    rp, ci = g.host_arrays()
    lo, hi = engine.ranges[engine.r]
    c0, c1 = engine.col_lo, engine.col_hi
    d = engine.d
    rng = np.random.RandomState(1234 + engine.rank)
    deg = np.diff(rp[lo:hi + 1])
    hubs = lo + np.nonzero(deg > 2048)[0]
    single = g.handle(device) if engine.world > 1 else None
    if single is not None:
        single.tune_hot_rows(d * 4)
    worst, checked, bit_identical, mean_ok = 0.0, 0, True, True
    cur_ref = X0[:, c0:c1].contiguous()
    for level in range(1, levels + 1):
        sums, means = engine.run_levels(X0, level)
        torch.cuda.synchronize()
        rows = np.unique(np.concatenate([rng.choice(np.arange(lo, hi), min(rows_per_level, hi - lo),
                                                    replace=False), hubs[:2000]]))
        S, M = refex_oracle.aggregate_rows_c(rows, rp, ci, cur_ref.cpu().numpy())
        got_s = sums[torch.as_tensor(rows - lo, device=device)].double().cpu().numpy()
        got_m = means[torch.as_tensor(rows - lo, device=device)].double().cpu().numpy()
        for got, ref in ((got_s, S), (got_m, M)):
            nz = ref != 0
            if nz.any():
                worst = max(worst, float(np.max(np.abs(got[nz] - ref[nz]) / np.abs(ref[nz]))))
            if (got[~nz] != 0).any():
                worst = float('inf')
        checked += int(rows.size)
        degs = torch.as_tensor(np.maximum(deg, 1), device=device, dtype=torch.float32)[:, None]
        if not torch.equal(means, sums / degs):
            mean_ok = False
        if single is not None:
            ref_out = single.aggregate(cur_ref)
            if not (torch.equal(sums, ref_out[lo:hi, :d]) and torch.equal(means, ref_out[lo:hi, d:])):
                bit_identical = False
            if engine.exchange == 'peer' and not torch.equal(engine.peers.replicas[level & 1],
                                                             ref_out[:, d:]):
                bit_identical = False
            cur_ref = ref_out[:, d:].contiguous()
            del ref_out
            # the replicas are views the peers overwrite in their next run_levels: nobody may
            # start the next level count before everybody has finished comparing
            torch.cuda.synchronize()
            dist.barrier()
        else:
            cur_ref = means.contiguous() if means.shape[0] == g.n else cur_ref
    out = {'max_rel_err': worst, 'tolerance': RTOL, 'rows_checked': checked,
           'levels_checked': levels,
           'oracle': 'float64 C restatement (oracle/refex_oracle.c) on sampled + hub rows, fed '
                     'with the level input the GPU used',
           'mean_is_correctly_rounded_sum_over_degree': mean_ok}
    if dist is not None:
        t = torch.tensor([worst if np.isfinite(worst) else 1e30, float(checked),
                          0.0 if bit_identical else 1.0, 0.0 if mean_ok else 1.0],
                         device=device, dtype=torch.float64)
        w = t.clone()
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
        s = t.clone()
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        out.update(max_rel_err=float(w[0]), rows_checked=int(s[1]),
                   mean_is_correctly_rounded_sum_over_degree=bool(float(w[3]) == 0.0),
                   bit_identical_to_single_gpu=bool(float(w[2]) == 0.0),
                   bit_identity='every rank re-ran the unsharded kernel on its column group: own '
                                'rows of every level' +
                                (' and the whole replica' if engine.exchange == 'peer' else ''))
    out['ok'] = bool(out['max_rel_err'] <= RTOL and
                     out.get('bit_identical_to_single_gpu', True) and
                     out['mean_is_correctly_rounded_sum_over_degree'])
    return out


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
    # pinned staging buffers of the e2e path should live on the GPU's own NUMA node
    numa_note = shard_mod.bind_host_thread_to_gpu(local_rank) if world > 1 else 'single process'
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

    # warm-up: measured-feedback balancing of the node ranges first (it runs >= 2 whole steps)
    history = engine.autobalance(X0, levels) if engine.R > 1 else []
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
    kern_ms = [ev[0].elapsed_time(ev[1]) for ev in level_events]
    wait_ms = [ev[1].elapsed_time(ev[2]) for ev in level_events if len(ev) > 2]
    kern_ms_avg = float(np.mean(kern_ms)) if kern_ms else float('nan')
    per_rank = None
    if dist is not None:
        t = torch.tensor([ms_total], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        stats = torch.zeros(world, 4, device=device, dtype=torch.float64)
        stats[rank] = torch.tensor([kern_ms_avg, float(np.mean(wait_ms)) if wait_ms else 0.0,
                                    float(engine.local_rows), float(engine.local_nnz)])
        dist.all_reduce(stats)
        per_rank = stats.cpu().numpy()
    ms_per_step = ms_total / args.steps
    value = nnz * d * levels / (ms_per_step * 1e-3)

    # dominant kernel: the gather-reduce launch of each level (events bracket one
    # gr_refex_aggregate[_bcast]_f32 call = gather kernel + the microsecond-scale hub fix-up)
    peak, peak_src = measured_peak_hbm()
    alg_bytes = algorithmic_bytes_per_level(engine.local_rows, engine.local_nnz, engine.d)
    nvlink_bytes = 0
    if engine.exchange == 'peer':        # the mean rows also go to the R - 1 peer replicas
        nvlink_bytes = engine.local_rows * engine.d * 4 * (engine.R - 1)
    achieved = alg_bytes / (kern_ms_avg * 1e-3) / 1e9
    traffic = NCU_TRAFFIC.get(args.workload) if world == 1 else None

    parallelism = 'single GPU' if world == 1 else \
        (f'{engine.C} column group(s) x {engine.R} node range(s), {engine.balance}, full replica '
         f'of the column group\'s input per GPU')
    exchange = {'none': 'none (no exchange step: ' +
                        ('single GPU)' if world == 1 else 'column groups are independent)'),
                'peer': 'fused: gather kernel stores mean rows into every replica of its column '
                        'group over NVLink-mapped peer pointers + flag barrier',
                'nccl': 'all-gather of the mean rows after the kernel'}[engine.exchange] + \
        (f' [{engine.exchange_note}]' if engine.exchange_note else '')
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.workload, n, nnz, d, levels),
        'sharding': {'parallelism': parallelism, 'exchange': exchange,
                     'column_groups': engine.C, 'node_ranges': engine.R,
                     'value_with_undirected_edge_convention': value / 2},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                     'frac': achieved / peak,
                     'traffic': traffic['bytes'] if traffic else None,
                     'traffic_source': (traffic['source'] + ' (ncu capture of the same kernel and '
                                        'workload, not re-measured in this run)') if traffic else None,
                     'peak_source': peak_src,
                     'kernel': 'refex_gather_kernel' if engine.exchange != 'peer'
                     else 'refex_gather_bcast_kernel', 'kernel_ms_avg': kern_ms_avg,
                     'algorithmic_bytes_per_launch': alg_bytes,
                     'nvlink_store_bytes_per_launch': nvlink_bytes,
                     'note': None if engine.d * 4 >= 256 else
                     f'rows of {engine.d * 4} bytes: random gathers below 256 bytes are bound by '
                     f'DRAM row activations, not bytes (DESIGN.md section 5)'},
        'clocks': clocks.summary(),
        'gpu_launches': launches,
    }
    if per_rank is not None:
        line['sharding']['per_rank'] = {
            'kernel_ms_per_level': [round(float(x), 4) for x in per_rank[:, 0]],
            'barrier_wait_ms_per_level': [round(float(x), 4) for x in per_rank[:, 1]],
            'rows': [int(x) for x in per_rank[:, 2]], 'arcs': [int(x) for x in per_rank[:, 3]],
            'nvlink_store_bytes_per_level': [int(x) * engine.d * 4 * (engine.R - 1)
                                             if engine.exchange == 'peer' else 0
                                             for x in per_rank[:, 2]]}
        line['sharding']['balancing_history'] = history

    # ---- parity of what was timed (outside the timed region) ----------------------------------
    if not args.no_parity:
        try:
            line['parity'] = verify_parity(engine, g, X0, levels, dist, device)
        except Exception as exc:
            line['parity'] = {'ok': False, 'error': repr(exc)}

    # ---- e2e: host buffers through the C-ABI, copies inside the timed region ------------------
    if not args.no_e2e:
        try:
            line['e2e'] = e2e_section(engine, X0, n, nnz, d, levels, dist, device, args)
            line['e2e']['host_numa'] = numa_note
        except Exception as exc:
            line['e2e'] = {'error': repr(exc)}
    engine.close()
    del engine
    torch.cuda.empty_cache()

    # ---- CPU baseline beside it (rank 0, N = 1 only) -----------------------------------------
    if world == 1 and not args.no_cpu_baseline:
        rp, ci = g.host_arrays()
        X_host = X0.cpu().numpy()
        rng = np.random.RandomState(0)
        ref = CpuReference(rp, ci, X_host, d)
        rate, k, arcs, dt = ref.sample(12.0, rng)
        fair, threads, farcs, fdt = cpu_fair_comparator(rp, ci, X_host, d, min(n, 1_000_000))
        line['cpu_baseline'] = {
            'value': rate, 'unit': UNIT, 'cores': 1, 'kind': ref.kind,
            'sample': f'{k} uniformly sampled nodes ({arcs} arcs) of the same graph/features '
                      f'through {ref.describe()}, {dt:.1f} s',
            'host_cores': os.cpu_count(),
            'fair_c_openmp_f64': {'value': fair, 'unit': UNIT, 'cores': threads,
                                  'sample': f'first {min(n, 1_000_000)} rows ({farcs} arcs), '
                                            f'{fdt:.2f} s'}}
        del ref

    # ---- the "next" rows of the path on the same graph (N = 1) ----------
    if world == 1 and not args.no_next:
        try:
            line['next'] = next_rows_section(g, d, device)
        except Exception as exc:
            line['next'] = {'error': repr(exc)}

    # ---- BASELINE.json configs[1] beside the headline (N = 1) ----------------------------------
    if world == 1 and args.workload == 'c3' and not args.no_c2:
        try:
            del g, X0
            torch.cuda.empty_cache()
            line['c2'] = c2_section(device)
        except Exception as exc:
            line['c2'] = {'error': repr(exc)}

    # ---- hot path B beside it (N = 1): RolX NMF, BASELINE.json configs[4] ---------------------
    if world == 1 and not args.no_nmf:
        try:
            torch.cuda.empty_cache()
            line['nmf'] = nmf_section(device)
        except Exception as exc:      # never lose the headline line to the secondary section
            line['nmf'] = {'error': repr(exc)}

    # ---- hot path B row-sharded over the ranks (N > 1) -------------------
    if world > 1 and not args.no_nmf:
        torch.cuda.empty_cache()
        try:      # never lose the headline line to the secondary section
            line['nmf_row_sharded'] = nmf_row_sharded_section(device, dist, rank, world)
        except Exception as exc:
            line['nmf_row_sharded'] = {'error': repr(exc)}

    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def e2e_section(engine, X0, n, nnz, d, levels, dist, device, args):
    """The same metric through the host-buffer C-ABI call (gr_refex_levels_host_f32 /
    gr_refex_levels_host_sharded_f32): every step copies this rank's columns of X0 from pinned
    host memory and copies its rows of EVERY level's result back; max over ranks."""
    lo, hi = engine.ranges[engine.r]
    Xh = torch.empty((n, d), dtype=torch.float32, pin_memory=True)
    Xh.copy_(X0)
    shape = (levels, hi - lo, 2 * engine.d) if engine.R == 1 else (levels, 2, hi - lo, engine.d)
    outh = torch.empty(shape, dtype=torch.float32, pin_memory=True)
    engine.run_levels_host(Xh, levels, outh)       # warm-up (allocates staging)
    e2e_steps = max(1, min(args.steps, 3))
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        engine.run_levels_host(Xh, levels, outh)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / e2e_steps
    # a node-range shard uploads only its own rows of X0 (the peers get them over NVLink)
    h2d = (n if engine.R == 1 else hi - lo) * engine.d * 4
    d2h = levels * (hi - lo) * 2 * engine.d * 4
    if dist is not None:
        t = torch.tensor([dt, float(h2d), float(d2h)], device=device, dtype=torch.float64)
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dt, h2d, d2h = float(mx[0]), int(t[1]), int(t[2])
    # spot check: the last level that came back equals the device-resident result
    sums, means = engine.run_levels(X0, levels)
    last = outh[levels - 1]
    got_s, got_m = (last[:, :engine.d], last[:, engine.d:]) if engine.R == 1 else (last[0], last[1])
    same = torch.equal(got_s.to(device), sums) and torch.equal(got_m.to(device), means)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    return {'value': nnz * d * levels / dt, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
            'd2h_bytes_per_step': d2h, 'ms_per_step': dt * 1e3, 'steps': e2e_steps,
            'matches_device_path': bool(same),
            'api': ('gr_refex_levels_host_f32' if engine.R == 1 else
                    'gr_refex_levels_host_sharded_f32') +
                   ' (pinned host X in, own rows of all levels out; bytes summed over ranks, '
                   'time = max over ranks)'}


def c2_section(device, steps=20):
    """BASELINE.json configs[1] (ER 1 M / 20 M, 32 features, 4 levels): schedule alpha and the
    reference-shaped schedule beta (input width doubles per level: 32 -> 64 -> 128 -> 256)."""
    from graphrole_b200 import shard as shard_mod
    g, d, levels = build_graph('c2', device)
    X0 = torch.rand(g.n, d, device=device, generator=torch.Generator(device=device).manual_seed(0))
    peak, _ = measured_peak_hbm()
    eng = shard_mod.ShardedRefex(g, d)
    for _ in range(3):
        eng.run_levels(X0, levels)
    events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.run_levels(X0, levels, events)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    kern = float(np.mean([a.elapsed_time(b) for a, b in events]))
    alg = algorithmic_bytes_per_level(g.n, g.nnz, d)
    out = {'config': workload_config('c2', g.n, g.nnz, d, levels), 'ms_per_step': ms,
           'value': g.nnz * d * levels / ms * 1e3, 'unit': UNIT,
           'roofline': {'kernel_ms_avg': kern, 'algorithmic_bytes_per_launch': alg,
                        'achieved': alg / kern / 1e6, 'frac': alg / kern / 1e6 / peak,
                        'note': 'the 128 MB feature matrix is about the size of L2: DRAM traffic '
                                'is below the algorithmic bytes'}}
    parity = verify_parity(eng, g, X0, levels, None, device)
    out['parity'] = parity
    # schedule beta: every level aggregates ALL columns the previous level produced
    h = g.handle(device)
    widths = [d * 2 ** l for l in range(levels)]
    bufs = [torch.empty((g.n, 2 * w), device=device) for w in widths]

    def beta():
        cur = X0
        for l in range(levels):
            cur = h.aggregate(cur, out=bufs[l])
    for _ in range(2):
        beta()
    e0.record()
    for _ in range(5):
        beta()
    e1.record()
    torch.cuda.synchronize()
    ms_b = e0.elapsed_time(e1) / 5
    out['schedule_beta'] = {'widths': widths, 'ms_per_step': ms_b,
                            'value': g.nnz * sum(widths) / ms_b * 1e3, 'unit': UNIT,
                            'alg_GBps': sum(algorithmic_bytes_per_level(g.n, g.nnz, w)
                                            for w in widths) / ms_b / 1e6}
    return out


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
    peak, _ = measured_peak_hbm()
    ms = timed(lambda: level0.device_features(g))
    # compulsory traffic of the triangle path: rowptr + colidx read by the orientation pass, the
    # oriented copy written and read, counters: ~ (8 n + 4 nnz) * 2 + 4 nnz/2 * 2 + 24 n bytes
    l0_bytes = (8 * g.n + 4 * g.nnz) * 2 + 4 * (g.nnz // 2) * 2 + 24 * g.n
    out['level0'] = {'ms': ms, 'arcs_per_s': g.nnz / ms * 1e3,
                     'columns': ['degree', 'internal_edges', 'external_edges'],
                     'path': 'triangle counting on the degree-oriented graph (undirected, '
                             'unweighted, no self loops)',
                     'roofline': {'bound': 'latency of dependent binary-search loads (L2), not a '
                                           'stream', 'compulsory_bytes': l0_bytes,
                                  'achieved': l0_bytes / ms / 1e6, 'peak': peak, 'unit': 'GB/s',
                                  'frac': l0_bytes / ms / 1e6 / peak}}
    feats = g.handle(device).aggregate(torch.rand(g.n, d, device=device))
    pruner = _native.Pruner(g.n, device)
    bins = torch.empty((feats.shape[1], g.n), dtype=torch.int32, device=device)
    ms = timed(lambda: pruner.bin_columns(feats, out=bins))
    # per key: transpose pass (4 B in, 4 out), radix sort of (key, row) pairs = 4 passes x
    # (8 B in + 8 B out), bin scatter (4 B in, 4 B out) = 80 bytes
    bin_bytes = 80 * feats.numel()
    out['vertical_log_binning'] = {'columns': feats.shape[1], 'ms': ms,
                                   'keys_per_s': feats.numel() / ms * 1e3,
                                   'roofline': {'bound': 'hbm (cub radix sort passes)',
                                                'algorithmic_bytes': bin_bytes,
                                                'achieved': bin_bytes / ms / 1e6, 'peak': peak,
                                                'unit': 'GB/s', 'frac': bin_bytes / ms / 1e6 / peak}}
    ms = timed(lambda: pruner.pairwise_gaps(bins), reps=2)
    f = bins.shape[0]
    # 2 fp32 ALU operations (|a - b|, max) per (pair, row); 148 SMs x 128 lanes per clock
    pair_rows = f * (f - 1) / 2 * g.n
    alu_peak = 148 * 128 * 1.965e9 / 2
    out['pairwise_gaps'] = {'columns': f, 'ms': ms, 'pair_rows_per_s': pair_rows / ms * 1e3,
                            'roofline': {'bound': 'fp32 alu (2 operations per pair and row)',
                                         'achieved': pair_rows / ms * 1e3, 'peak': alu_peak,
                                         'unit': 'pair*rows/s',
                                         'frac': pair_rows / ms * 1e3 / alu_peak}}
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


def nmf_section(device, n=10_000_000, f=512, ranks=(2, 4, 5, 8, 16, 32), iters=10):
    """Hot path B on C5 (X = 10M x 512 fp32, synthetic U[0,1)), tcgen05 path:
    * ms per multiplicative-update iteration with tol = 0 (no convergence pass in the timed
      region); algorithmic bytes n*f*4 + 2*n*r*4;
    * the same with sklearn's defaults (tol = 1e-4, dense-residual check every 10 iterations):
      what get_nmf_decomposition actually runs (the checks run W H on tcgen05 as well), and one
      check alone; ranks 2 and 5 stand for the n_roles of the reference's default grid (2..8) that
      are not a multiple of 4: the library runs them on zero-padded factors;
    * the RolX epilogue at this size: quantiser, description-length cost and the whole
      (n_roles, n_bits) model-selection grid on device-resident factors."""
    from graphrole_b200 import _native
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
        # sklearn's defaults: 20 iterations = two convergence checks (+ the error at init)
        e0.record()
        n_it, err = solver.update(X, W, H, max_iter=20, tol=1e-30, check_every=10)
        e1.record()
        torch.cuda.synchronize()
        ms_checked = e0.elapsed_time(e1) / max(n_it, 1)
        e0.record()
        solver.update(X, W, H, max_iter=0, tol=1e-30)      # the error at init alone = one check
        e1.record()
        torch.cuda.synchronize()
        ms_check = e0.elapsed_time(e1)
        alg = n * f * 4 + 2 * n * r * 4
        flops = 4 * n * f * r + 4 * n * r * r + 2 * r * r * f
        rows.append({'r': r, 'path': solver.last_path, 'ms_per_iter': ms,
                     'alg_GBps': alg / ms / 1e6, 'frac_of_hbm_peak': alg / ms / 1e6 / peak,
                     'alg_TFLOPs': flops / ms / 1e9,
                     'ms_per_iter_with_convergence_checks': ms_checked,
                     'iterations_with_checks': n_it, 'ms_per_convergence_check': ms_check,
                     'check_GBps_of_X': n * f * 4 / ms_check / 1e6})
        solver.close()
        del W, H
    out = {'workload': f'X {n}x{f} fp32 U[0,1), shared random init, {iters} iterations',
           'dtype': 'tf32 MMA / fp32 accumulate', 'per_rank': rows}
    # roofline objects of the two kernels of the path (largest rank timed): algorithmic bytes over
    # the CUDA-event time of the whole iteration / check, against the measured HBM peak
    last = rows[-1]
    r_last = last['r']
    out['roofline'] = {
        'iteration': {
            'kernel': 'nmf_fused_tc_kernel<2> (+ H H^T, partial reduction, H update: 3 small launches)',
            'bound': 'hbm', 'r': r_last, 'achieved': last['alg_GBps'], 'peak': peak, 'unit': 'GB/s',
            'frac': last['frac_of_hbm_peak'],
            'algorithmic_bytes_per_iteration': n * f * 4 + 2 * n * r_last * 4,
            'traffic': None,
            'traffic_source': 'profiles/r2_ncu_nmf_tc_raw.csv (n = 4 M, r = 32: dram read 8.77 GB '
                              'for 8.19 + 0.51 GB algorithmic; not re-measured in this run)',
            'note': 'co-bound by the SM shared-memory data pipe (DESIGN.md section 6)'},
        'convergence_check': {
            'kernel': 'nmf_error_tc_kernel', 'bound': 'hbm', 'r': r_last,
            'achieved': (n * f * 4 + n * r_last * 4) / last['ms_per_convergence_check'] / 1e6,
            'peak': peak, 'unit': 'GB/s',
            'frac': (n * f * 4 + n * r_last * 4) / last['ms_per_convergence_check'] / 1e6 / peak,
            'algorithmic_bytes_per_check': n * f * 4 + n * r_last * 4,
            'traffic': None,
            'traffic_source': 'profiles/r2_ncu_nmf_error_tc_raw.csv (n = 4 M, r = 32: dram read '
                              '8.71 GB for 8.70 GB algorithmic at 6.83 TB/s)',
            'note': 'read-only stream: above the copy peak of MEASURED_PEAKS.json; the time '
                    'includes the D2H of the partials and the stream synchronisation'}}
    try:
        out['rolx_epilogue'] = rolx_epilogue_section(X, device)
    except Exception as exc:
        out['rolx_epilogue'] = {'error': repr(exc)}
    try:      # sklearn MU (the reference's solver) on a row sample, all BLAS threads
        from sklearn.decomposition import _nmf as sk
        m = 200_000
        Xc = X[:m].double().cpu().numpy()
        rng = np.random.RandomState(0)
        W0, H0 = rng.rand(m, 32) + 0.1, rng.rand(32, f) + 0.1
        t0 = time.perf_counter()
        W_sk, H_sk, _ = sk._fit_multiplicative_update(Xc, W0.copy(), H0.copy(), 'frobenius',
                                                      max_iter=2, tol=0)
        dt = (time.perf_counter() - t0) / 2
        out['cpu_sklearn_f64'] = {'r': 32, 'rows': m, 's_per_iter_sample': dt,
                                  's_per_iter_scaled_to_n': dt * n / m,
                                  'cores': os.cpu_count()}
        # parity of what was timed: the same rows, the same start, the same two iterations through
        # the tcgen05 kernels against scikit-learn's own loop (TF32 tolerance of the tests: 1e-2
        # of the factor's largest entry; error 1e-3)
        Wg, Hg, _, err_g = factor.nmf_mu(
            X[:m], torch.as_tensor(W0, dtype=torch.float32, device=device),
            torch.as_tensor(H0, dtype=torch.float32, device=device), max_iter=2, tol=0)
        dw = float(np.abs(Wg.double().cpu().numpy() - W_sk).max() / np.abs(W_sk).max())
        dh = float(np.abs(Hg.double().cpu().numpy() - H_sk).max() / np.abs(H_sk).max())
        err_sk = float(np.linalg.norm(Xc - W_sk @ H_sk))
        out['parity'] = {'against': 'sklearn _fit_multiplicative_update (float64), same rows / '
                                    'start / iterations', 'rows': m, 'r': 32, 'iterations': 2,
                         'path': factor.last_path, 'rel_dW': dw, 'rel_dH': dh,
                         'error': err_g, 'error_sklearn': err_sk, 'tolerance': 1e-2,
                         'ok': bool(dw < 1e-2 and dh < 1e-2 and
                                    abs(err_g - err_sk) <= 1e-3 * err_sk and
                                    factor.last_path == 'tcgen05')}
    except Exception as exc:
        out['cpu_sklearn_f64'] = {'error': repr(exc)}
    return out


def nmf_row_sharded_section(device, dist, rank, world, n=10_000_000, f=512, ranks=(8, 32),
                            iters=20):
    """C5 with its rows split over the ranks (strong scaling of hot path B): H replicated, one
    NCCL all-reduce of [W^T X | W^T W] per iteration (roles/sharded.py).  Every rank reaches the
    collectives together or not at all: set-up is voted on first."""
    from graphrole_b200.roles.sharded import RowShardedNmf, row_shard
    lo, hi = row_shard(n, world, rank)
    out = {'workload': f'X {n}x{f} fp32 U[0,1) split by rows over {world} ranks, H replicated',
           'rows_per_rank': hi - lo, 'per_rank': []}
    X = solver = None
    ok = torch.ones(1, device=device)
    try:
        gen = torch.Generator(device=device).manual_seed(100 + rank)
        X = torch.rand(hi - lo, f, device=device, generator=gen)
    except Exception as exc:
        out['error'] = repr(exc)
        ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if not bool(ok.item()):
        out.setdefault('error', 'set-up failed on another rank')
        return out
    for r in ranks:
        ok.fill_(1)
        try:
            W = torch.rand(hi - lo, r, device=device, generator=gen) + 0.1
            H = torch.rand(r, f, device=device,
                           generator=torch.Generator(device=device).manual_seed(99)) + 0.1
            solver = RowShardedNmf(hi - lo, f, r, device)
        except Exception as exc:
            out['error'] = repr(exc)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not bool(ok.item()):
            break
        solver.fit(X, W, H, max_iter=3, tol=0)
        row = {'r': r, 'path': solver.backend.last_path,
               'allreduce_bytes_per_iteration': solver.allreduce_bytes_per_iteration}
        for key, kw in (('ms_per_iter', dict(tol=0)),
                        ('ms_per_iter_with_convergence_checks', dict(tol=1e-30, check_every=10))):
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n_it, _ = solver.fit(X, W, H, max_iter=iters, **kw)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / max(n_it, 1)], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            row[key] = float(t.item())
        # parity of what was timed: H is replicated, so every rank must hold the same bits
        Hs = [torch.empty_like(H) for _ in range(world)]
        dist.all_gather(Hs, H)
        row['H_identical_on_all_ranks'] = all(torch.equal(h, Hs[0]) for h in Hs)
        row['H_finite_nonnegative'] = bool(torch.isfinite(H).all() and (H >= 0).all())
        out['per_rank'].append(row)
        solver.close()
        del W, H
    return out


def rolx_epilogue_section(X, device, n_roles=8):
    """SURVEY.md section 8f #4 at C5 scale, factors resident in HBM: bind (sort + prefix sums)
    and encode of the n x r node-role factor for 2^1..2^8 bins, one description-length error cost
    (a pass over X), and the wall time of a whole model-selection grid."""
    from graphrole_b200 import _native
    from graphrole_b200.roles.extract import DeviceModelGrid
    n, f = X.shape
    out = {}
    gen = torch.Generator(device=device).manual_seed(1)
    W = torch.rand(n, n_roles, device=device, generator=gen) ** 2
    H = torch.rand(n_roles, f, device=device, generator=gen)

    def wall(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3, r

    q = _native.Quantizer(W.numel(), device)
    ms, _ = wall(lambda: q.bind(W))
    out['bind_ms'] = ms
    enc = {}
    G8 = None
    for bits in (1, 2, 4, 6, 8):
        ms, (G, info) = wall(lambda: q.encode(2 ** bits))
        enc[str(2 ** bits)] = {'ms': ms, 'lloyd_iterations': info['n_iter'],
                               'distinct': info['n_distinct']}
        G8 = G
    out['encode_node_role_factor'] = {'entries': W.numel(), 'by_bins': enc}
    q.close()
    ms, cost = wall(lambda: _native.mdl_error_cost(X, G8, H))
    out['error_cost'] = {'ms': ms, 'entries': n * f,
                         'GBps_of_X': n * f * 4 / ms / 1e6, 'value': cost}
    del W, H, G8
    torch.cuda.empty_cache()
    # one-time initialisation of the dense linear-algebra libraries behind torch.linalg (cuSOLVER /
    # MAGMA handles and workspaces: ~2 s in a fresh process) is not part of the grid
    t0 = time.perf_counter()
    from graphrole_b200.roles import factor
    factor.nndsvda_init(torch.rand(4096, 64, device=device), 4, seed=0)
    torch.cuda.synchronize()
    out['linalg_warmup_s'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    grid = DeviceModelGrid.from_device(X)
    grid.timed = True
    cells = 0
    for roles in (2, 4, 8):
        for bits in range(1, 9):
            grid.costs(roles, bits)
            cells += 1
    torch.cuda.synchronize()
    out['model_selection_grid'] = {'cells': cells, 'n_roles': [2, 4, 8], 'n_bits': [1, 8],
                                   'nmf_fits': grid.n_fits, 'wall_s': time.perf_counter() - t0,
                                   'phase_s': {k: round(v, 3) for k, v in grid.timings_s.items()},
                                   'nmf_iterations': dict(grid.nmf_iterations)}
    grid.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c3', choices=list(WORKLOADS))
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-nmf', action='store_true')
    ap.add_argument('--no-next', action='store_true')
    ap.add_argument('--no-c2', action='store_true')
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
