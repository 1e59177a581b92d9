#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_prune_level0_gpu.py -q -m gpu -x > gpurun_out/r2c39_pytest_level0.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2c39_pytest_level0.log
timeout 300 python - > gpurun_out/r2c39_level0.txt 2>&1 <<'P'
import torch, json, time, sys, os
sys.path.insert(0, os.getcwd())
from graphrole_b200.graph.generators import barabasi_albert_csr, erdos_renyi_csr
from graphrole_b200.graph import level0
for name, g in (('BA 10M m=20', barabasi_albert_csr(10_000_000, 20, seed=0, device='cuda:0')),
                ('ER 1M 20M edges', erdos_renyi_csr(1_000_000, 20_000_000, seed=0, device='cuda:0'))):
    for env in ('', '1'):
        ts = []
        for rep in range(4):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); out = level0.device_features(g); e1.record(); torch.cuda.synchronize()
            ts.append(round(e0.elapsed_time(e1), 2))
        print(json.dumps({'graph': name, 'level0_ms': ts}), flush=True)
        break
P
cat gpurun_out/r2c39_level0.txt | tail -4
