#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_prune_level0_gpu.py -q -m gpu -x -k "level0" > gpurun_out/r2c36_pytest_level0.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2c36_pytest_level0.log
timeout 300 python - > gpurun_out/r2c36_level0.txt 2>&1 <<'P'
import torch, json, time, sys, os
sys.path.insert(0, os.getcwd())
from graphrole_b200.graph.generators import barabasi_albert_csr
from graphrole_b200.graph import level0
g = barabasi_albert_csr(10_000_000, 20, seed=0, device='cuda:0')
for rep in range(4):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); out = level0.device_features(g); e1.record(); torch.cuda.synchronize()
    print(json.dumps({'rep': rep, 'level0_ms': round(e0.elapsed_time(e1), 2)}), flush=True)
P
cat gpurun_out/r2c36_level0.txt | tail -4
