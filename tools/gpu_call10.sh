#!/bin/bash
mkdir -p gpurun_out
GR_NMF_TC_DEBUG=32 timeout 120 python tools/debug_nmf_tc.py 1000,512,32 28433,512,32 2>&1 | cut -c1-120 | tail -3
GR_NMF_TC_DEBUG=32 timeout 120 python tools/bench_nmf.py --ranks 32 --iters 10 --paths tcgen05 2>&1 | cut -c1-200
GR_NMF_TC_DEBUG=32 GR_NMF_TRACE=1 GR_NMF_TRACE_FIRST=20 timeout 120 python tools/bench_nmf.py --n 2000000 --ranks 32 --iters 1 --paths tcgen05 > gpurun_out/nmf_pair_trace32.log 2>&1
grep "^blk" gpurun_out/nmf_pair_trace32.log | tail -8 | cut -c1-160
