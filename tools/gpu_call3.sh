#!/bin/bash
# session 3, call 1: new pruning / level-0 kernels (parity + timing), the K-major 32B-atom swizzle
# experiment of the NMF kernel, then the whole GPU suite.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python -m pytest tests/test_prune_level0_gpu.py -m gpu -x -q > gpurun_out/pytest_next.log 2>&1; echo "pytest-next rc=$? t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/pytest_next.log
tail -15 gpurun_out/pytest_next.log
GR_NMF_TC_DEBUG=16 timeout 300 python tools/debug_nmf_tc.py 64,128,32 1000,512,32 9489,512,32 > gpurun_out/nmf_dbg16.log 2>&1; echo "dbg16 rc=$? t=$(( $(date +%s) - t0 ))s"
cat gpurun_out/nmf_dbg16.log | tail -8
timeout 600 python tools/bench_next.py > gpurun_out/bench_next.log 2>&1; echo "bench-next rc=$? t=$(( $(date +%s) - t0 ))s"
tail -6 gpurun_out/bench_next.log
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_prune_level0_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest-all rc=$? t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
