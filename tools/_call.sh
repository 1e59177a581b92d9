#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_prune_level0_gpu.py -q -m gpu -x -k "triangle or level0" > gpurun_out/r2c43_pytest_level0.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c43_pytest_level0.log
