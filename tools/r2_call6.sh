#!/bin/bash
# round 2, call 6 (1 GPU): whole GPU test suite, NMF timing (error pass), compute-sanitizer
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu -s > gpurun_out/r2c6_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; grep -E "pair kernel vs|max relative|passed|failed|Error" gpurun_out/r2c6_pytest_gpu.log | tail -16
timeout 600 python - > gpurun_out/r2c6_nmf_checks.txt 2>&1 <<'PY'
import json, sys, torch
sys.path.insert(0, '.')
import bench
from graphrole_b200.roles import factor
dev = torch.device('cuda', 0)
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
    e0.record(); err2 = factor.nmf_error(X, W, H); e1.record(); torch.cuda.synchronize()
    print(json.dumps({'r': r, 'ms_per_iter': plain, 'ms_per_iter_with_checks': checked, 'iters': it,
                      'error_pass_ms': e0.elapsed_time(e1), 'err': err, 'err2': err2}))
    s.close()
PY
echo "nmf checks rc=$?"; cat gpurun_out/r2c6_nmf_checks.txt | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c6_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; grep -E "ERROR SUMMARY|smoke:|Invalid|Error" gpurun_out/r2c6_sanitizer_memcheck_smoke.log | tail -8
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c6_sanitizer_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; grep -E "RACECHECK SUMMARY|smoke:|hazard|Error" gpurun_out/r2c6_sanitizer_racecheck_smoke.log | tail -8
