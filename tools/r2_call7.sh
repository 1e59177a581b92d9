#!/bin/bash
# round 2, call 7 (1 GPU): whole GPU suite (no -x), level-0 fast path timing, NMF error pass
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -s > gpurun_out/r2c7_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; grep -E "pair kernel vs|max relative|passed|failed|^FAILED|^ERROR" gpurun_out/r2c7_pytest_gpu.log | tail -24
timeout 600 python - > gpurun_out/r2c7_level0_nmf.txt 2>&1 <<'PY'
import json, os, sys, time, torch
sys.path.insert(0, '.')
from graphrole_b200.graph.generators import barabasi_albert_csr
from graphrole_b200.graph import level0
from graphrole_b200.roles import factor
from graphrole_b200.features.device import DeviceRecursiveFeatureExtractor
dev = torch.device('cuda', 0)
g = barabasi_albert_csr(10_000_000, 20, seed=0, device=dev)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
fast = timed(lambda: level0.device_features(g))
a = level0.device_features(g)
os.environ['GR_LEVEL0_GENERAL'] = '1'
general = timed(lambda: level0.device_features(g), reps=1)
b = level0.device_features(g)
del os.environ['GR_LEVEL0_GENERAL']
print(json.dumps({'level0_fast_ms': fast, 'level0_general_ms': general,
                  'equal': all(bool(torch.equal(a[k], b[k])) for k in a)}))
t0 = time.perf_counter()
rfe = DeviceRecursiveFeatureExtractor(g, device=dev)
names, values = rfe.extract_features_device(); torch.cuda.synchronize()
print(json.dumps({'extract_features_wall_s': time.perf_counter() - t0, 'generations': rfe.generation_count,
                  'features': len(names), 'kernel_ms': {k: round(v, 2) for k, v in rfe.timings_ms.items()}}))
del g, rfe, values, a, b
torch.cuda.empty_cache()
n, f = 10_000_000, 512
gen = torch.Generator(device=dev).manual_seed(0)
X = torch.rand(n, f, device=dev, generator=gen)
for r in (4, 32):
    W = torch.rand(n, r, device=dev, generator=gen) + 0.1
    H = torch.rand(r, f, device=dev, generator=gen) + 0.1
    s = factor.NmfSolver(n, f, r, dev)
    s.update(X, W, H, max_iter=3, tol=0, want_error=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.update(X, W, H, max_iter=20, tol=0, want_error=False); e1.record(); torch.cuda.synchronize()
    plain = e0.elapsed_time(e1) / 20
    e0.record(); it, err = s.update(X, W, H, max_iter=20, tol=1e-30, check_every=10); e1.record(); torch.cuda.synchronize()
    checked = e0.elapsed_time(e1) / max(it, 1)
    factor.nmf_error(X, W, H)
    e0.record(); err2 = factor.nmf_error(X, W, H); e1.record(); torch.cuda.synchronize()
    print(json.dumps({'r': r, 'ms_per_iter': plain, 'ms_per_iter_with_checks': checked, 'iters': it,
                      'error_pass_ms': e0.elapsed_time(e1), 'err': err, 'err2': err2}))
    s.close()
PY
echo "level0/nmf rc=$?"; tail -5 gpurun_out/r2c7_level0_nmf.txt | cut -c1-500
