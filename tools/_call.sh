#!/bin/bash
mkdir -p gpurun_out
timeout 60 python tools/bench_encode.py > gpurun_out/r2c51_bench_encode.txt 2>&1; head -3 gpurun_out/r2c51_bench_encode.txt | cut -c1-200; python - <<'P'
import torch, time, sys, os
sys.path.insert(0, os.getcwd())
from graphrole_b200 import _native
W = torch.rand(10_000_000, 8, device='cuda:0') ** 2
q = _native.Quantizer(W.numel(), 'cuda:0')
q.bind(W); torch.cuda.synchronize()
t0 = time.perf_counter(); q.bind(W); torch.cuda.synchronize(); print('second bind ms', round((time.perf_counter() - t0) * 1e3, 2))
P
