#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_rolx_gpu.py -q -m gpu -x -k "seeding or golden or one_bind or float32_storage or strided" > gpurun_out/r2c50_pytest_rolx.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|Mismatched" gpurun_out/r2c50_pytest_rolx.log | head -8
