#!/bin/bash
# round 2, call 8 (1 GPU): NMF parity after the tf32 rounding of W, ncu of the error pass
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nmf_gpu.py -q -s > gpurun_out/r2c8_nmf.log 2>&1; echo "nmf rc=$?"; grep -E "^.*pair kernel vs sklearn \[|passed|failed|^FAILED" gpurun_out/r2c8_nmf.log | grep -v "print(" | tail -16
timeout 600 python -m pytest tests/test_prune_level0_gpu.py -q > gpurun_out/r2c8_prune.log 2>&1; echo "prune rc=$?"; tail -2 gpurun_out/r2c8_prune.log
cat > /tmp/nmf_err.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from graphrole_b200.roles import factor
dev = torch.device('cuda', 0)
n, f, r = 2_000_000, 512, 32
X = torch.rand(n, f, device=dev); W = torch.rand(n, r, device=dev); H = torch.rand(r, f, device=dev)
print(factor.nmf_error(X, W, H))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nmf_error -c 1 -o gpurun_out/r2c8_ncu_nmf_error -f python /tmp/nmf_err.py > gpurun_out/r2c8_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2c8_ncu.log
