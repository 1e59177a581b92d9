#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nmf_gpu.py tests/test_rolx_gpu.py -q -m gpu -k "nndsvda or get_nmf or role_extractor or rolx or select or grid" > gpurun_out/r2c25_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2c25_pytest.log | head
timeout 300 python tools/bench_init.py > gpurun_out/r2c25_init.txt 2>&1; cat gpurun_out/r2c25_init.txt | tail -4
