#!/bin/bash
# round 2, call 2: new RolX-epilogue kernels + BASELINE-size parity tests
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_rolx_gpu.py -x -q > gpurun_out/r2c2_rolx.log 2>&1; echo "rolx rc=$?"; tail -25 gpurun_out/r2c2_rolx.log
timeout 1500 python -m pytest tests/test_refex_gpu.py -x -q -s -k "config2 or config3 or exact_ties" > gpurun_out/r2c2_refex_size.log 2>&1; echo "refex size rc=$?"; grep -E "max relative|passed|failed|Error" gpurun_out/r2c2_refex_size.log | tail
