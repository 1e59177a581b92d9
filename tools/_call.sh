#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rolx_gpu.py -q -m gpu > gpurun_out/r2c48_pytest_rolx.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|Mismatched" gpurun_out/r2c48_pytest_rolx.log | head -12
