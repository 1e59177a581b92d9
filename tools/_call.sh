#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2c22_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c22_pytest_gpu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nmf_error_tc -s 1 -c 1 -o gpurun_out/r2c22_ncu_nmf_error_tc -f python tools/bench_nmf.py --n 4000000 --ranks 32 --paths tcgen05 --iters 2 > gpurun_out/r2c22_ncu.log 2>&1; echo "ncu rc=$?"
pick() { grep -o '"r": [0-9]*.*' "$1" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads('{' + l.strip())
    print({k: d[k] for k in ('r', 'ms_per_iter', 'ms_per_check', 'ms_per_iter_with_checks', 'iters_with_checks', 'error_at_check', 'error_pass')})
"; }
B="timeout 300 python tools/bench_nmf.py --paths tcgen05 --iters 20"
$B --ranks 4,8,16,32 > gpurun_out/r2c22_nmf_tc_error.txt 2>&1; echo "tc error pass:"; pick gpurun_out/r2c22_nmf_tc_error.txt
GR_NMF_ERROR_FFMA=1 $B --ranks 4,8,16,32 > gpurun_out/r2c22_nmf_ffma_error.txt 2>&1; echo "ffma error pass:"; pick gpurun_out/r2c22_nmf_ffma_error.txt
