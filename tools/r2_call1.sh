#!/bin/bash
# round 2, call 1: refex tests after the __fdiv_rn change; column-group sweep; pairwise tiles
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_refex_gpu.py -x -q > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c1_pytest.log
timeout 900 python tools/exp_colsplit.py > gpurun_out/r2c1_colsplit.txt 2> gpurun_out/r2c1_colsplit.err; echo "colsplit rc=$?"; tail -3 gpurun_out/r2c1_colsplit.err
timeout 300 ./tools/exp_pairwise > gpurun_out/r2c1_pairwise.txt 2>&1; echo "pairwise rc=$?"; tail -8 gpurun_out/r2c1_pairwise.txt
