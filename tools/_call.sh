#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nmf_gpu.py tests/test_rolx_gpu.py -q -m gpu > gpurun_out/r2c23_pytest_nmf_rolx.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2c23_pytest_nmf_rolx.log | head -30
pick() { grep -o '"r": [0-9]*.*' "$1" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads('{' + l.strip())
    print({k: d[k] for k in ('r', 'ms_per_iter', 'ms_per_check', 'ms_per_iter_with_checks', 'iters_with_checks', 'error_pass')}, d.get('path'))
"; }
timeout 300 python tools/bench_nmf.py --paths tcgen05 --iters 20 --ranks 2,3,5,7 > gpurun_out/r2c23_nmf_odd_ranks.txt 2>&1; echo "odd ranks (padded):"; pick gpurun_out/r2c23_nmf_odd_ranks.txt; grep -o '"path": "[a-z0-9]*"' gpurun_out/r2c23_nmf_odd_ranks.txt | tr '\n' ' '
GR_NMF_NO_RANK_PADDING=1 timeout 300 python tools/bench_nmf.py --paths tcgen05 --iters 5 --ranks 2,7 > gpurun_out/r2c23_nmf_odd_ranks_ffma.txt 2>&1; echo "odd ranks (no padding -> ffma):"; pick gpurun_out/r2c23_nmf_odd_ranks_ffma.txt;  grep -o '"path": "[a-z0-9]*"' gpurun_out/r2c23_nmf_odd_ranks_ffma.txt | tr '\n' ' '
