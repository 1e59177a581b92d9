#!/bin/bash
# NMF CTA-pair kernel: parity vs the FFMA path, NMF GPU tests, C5 timing, timeline
mkdir -p gpurun_out
timeout 300 python tools/debug_nmf_tc.py > gpurun_out/nmf_pair_parity.log 2>&1; echo "parity rc=$?"
cat gpurun_out/nmf_pair_parity.log | cut -c1-200 | tail -14
timeout 600 python -m pytest tests/test_nmf_gpu.py -m gpu -x -q > gpurun_out/pytest_nmf.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_nmf.log
timeout 300 python tools/bench_nmf.py --ranks 32,8 --iters 10 --paths tcgen05 > gpurun_out/nmf_pair_bench.log 2>&1
cat gpurun_out/nmf_pair_bench.log | cut -c1-220
GR_NMF_NO_CLUSTER=1 timeout 300 python tools/bench_nmf.py --ranks 32 --iters 10 --paths tcgen05 2>&1 | cut -c1-220
GR_NMF_TRACE=1 GR_NMF_TRACE_FIRST=20 timeout 300 python tools/bench_nmf.py --n 2000000 --ranks 32 --iters 1 --paths tcgen05 > gpurun_out/nmf_pair_trace.log 2>&1
grep "^blk" gpurun_out/nmf_pair_trace.log | tail -24 | cut -c1-160
