#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r2c14_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/r2c14_pytest_gpu.log; grep -E "^FAILED|^ERROR" gpurun_out/r2c14_pytest_gpu.log | head
