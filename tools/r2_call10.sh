#!/bin/bash
# round 2, call 10 (1 GPU): tiled error pass (tests + timing), ncu launch list + full capture of the gather
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nmf_gpu.py tests/test_prune_level0_gpu.py -q > gpurun_out/r2c10_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2c10_tests.log
timeout 600 python - > gpurun_out/r2c10_nmf_checks.txt 2>&1 <<'PY'
import json, sys, torch
sys.path.insert(0, '.')
from graphrole_b200.roles import factor
dev = torch.device('cuda', 0)
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
echo "nmf checks rc=$?"; cat gpurun_out/r2c10_nmf_checks.txt | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c10_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-next --no-nmf --no-c2 --no-parity > gpurun_out/r2c10_launches_bench.log 2>&1; echo "launch list rc=$?"; tail -1 gpurun_out/r2c10_launches_bench.log | cut -c1-200
cat > /tmp/one_level.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from graphrole_b200.graph.generators import barabasi_albert_csr
g = barabasi_albert_csr(10_000_000, 20, seed=0, device='cuda:0')
X = torch.rand(g.n, 64, device='cuda:0')
h = g.handle('cuda:0')
out = h.aggregate(X); out = h.aggregate(X, out=out)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:refex_gather_kernel -s 1 -c 1 -o gpurun_out/r2c10_ncu_gather_d64 -f python /tmp/one_level.py > gpurun_out/r2c10_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2c10_ncu.log
