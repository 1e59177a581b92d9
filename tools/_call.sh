#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_prune_level0_gpu.py -q > gpurun_out/r2c15_prune.log 2>&1; echo "prune rc=$?"; tail -3 gpurun_out/r2c15_prune.log
timeout 600 python - > gpurun_out/r2c15_pairwise.txt 2>&1 <<'PY'
import json, os, sys, torch
sys.path.insert(0, '.')
from graphrole_b200 import _native
dev = torch.device('cuda', 0)
n = 10_000_000
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
p = _native.Pruner(n, dev)
for d in (32, 64, 128, 256):
    bins = torch.randint(0, 24, (d, n), device=dev, dtype=torch.int32)
    row = {'columns': d, 'ms_default': round(timed(lambda: p.pairwise_gaps(bins)), 3)}
    os.environ['GR_PRUNE_TILE4'] = '1'
    row['ms_tile4'] = round(timed(lambda: p.pairwise_gaps(bins)), 3)
    del os.environ['GR_PRUNE_TILE4']
    print(json.dumps(row), flush=True)
    del bins
PY
echo "pairwise rc=$?"; cat gpurun_out/r2c15_pairwise.txt
