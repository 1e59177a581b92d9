#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rolx_gpu.py -q -m gpu -x > gpurun_out/r2c32_pytest_rolx.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/r2c32_pytest_rolx.log | head -20
timeout 300 python tools/bench_encode.py > gpurun_out/r2c32_bench_encode.txt 2>&1; cat gpurun_out/r2c32_bench_encode.txt | tail -10
