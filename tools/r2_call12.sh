#!/bin/bash
# round 2, call 12 (1 GPU): NMF after the epilogue change, 128-thread error pass, pool retention
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nmf_gpu.py tests/test_prune_level0_gpu.py tests/test_rolx_gpu.py -q > gpurun_out/r2c12_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c12_tests.log
timeout 600 python - > gpurun_out/r2c12_timing.txt 2>&1 <<'PY'
import json, sys, torch
sys.path.insert(0, '.')
from graphrole_b200.roles import factor
from graphrole_b200.graph.generators import barabasi_albert_csr
from graphrole_b200.graph import level0
dev = torch.device('cuda', 0)
g = barabasi_albert_csr(10_000_000, 20, seed=0, device=dev)
ts = []
for _ in range(4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); level0.device_features(g); e1.record(); torch.cuda.synchronize()
    ts.append(round(e0.elapsed_time(e1), 2))
print(json.dumps({'level0_ms_call_by_call_with_syncs_between': ts}))
del g
torch.cuda.empty_cache()
n, f = 10_000_000, 512
gen = torch.Generator(device=dev).manual_seed(0)
X = torch.rand(n, f, device=dev, generator=gen)
for r in (4, 8, 16, 32):
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
    print(json.dumps({'r': r, 'ms_per_iter': plain, 'ms_per_iter_with_checks': checked, 'iters': it,
                      'ms_per_check': (checked - plain) * it / 3, 'err': err}))
    s.close()
PY
echo "timing rc=$?"; cat gpurun_out/r2c12_timing.txt | tail -6
