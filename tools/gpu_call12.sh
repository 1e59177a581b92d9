#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/debug_nmf_tc.py 2>&1 | cut -c1-120 | tail -9
GR_NMF_NO_CLUSTER=1 timeout 120 python tools/debug_nmf_tc.py 1000,512,32 5000,96,12 2>&1 | cut -c1-120 | tail -2
timeout 300 python -m pytest tests/test_nmf_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/bench_nmf.py --ranks 32,8 --iters 10 --paths tcgen05 2>&1 | cut -c1-200
GR_NMF_NO_CLUSTER=1 timeout 120 python tools/bench_nmf.py --ranks 32 --iters 10 --paths tcgen05 2>&1 | cut -c1-200
GR_NMF_TRACE=1 GR_NMF_TRACE_FIRST=20 timeout 120 python tools/bench_nmf.py --n 2000000 --ranks 32 --iters 1 --paths tcgen05 > gpurun_out/nmf_pair_trace_e16.log 2>&1
grep "^blk" gpurun_out/nmf_pair_trace_e16.log | tail -8 | cut -c1-160
