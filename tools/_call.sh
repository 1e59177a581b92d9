#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_prune_level0_gpu.py -q -m gpu -x > gpurun_out/r2c31_pytest_prune.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/r2c31_pytest_prune.log | head -20
timeout 300 python tools/bench_next.py > gpurun_out/r2c31_bench_next.txt 2>&1; cat gpurun_out/r2c31_bench_next.txt | tail -4
