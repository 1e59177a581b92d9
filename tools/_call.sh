#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nmf_gpu.py -q -m gpu -k "row_sharded or tensor_core_path or any_feature" > gpurun_out/r2c26_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/r2c26_pytest.log | head -30
timeout 300 python tools/check_sharded_nmf.py --n 200000 --bench --bench-n 10000000 > gpurun_out/r2c26_check_nmf_n1.txt 2>&1; echo "check rc=$?"; tail -8 gpurun_out/r2c26_check_nmf_n1.txt | cut -c1-400
