#!/bin/bash
mkdir -p gpurun_out
echo "== CL=1 default rings"; GR_NMF_NO_CLUSTER=1 timeout 120 python tools/bench_nmf.py --ranks 4,8,16,32 --iters 10 --paths tcgen05 2>&1 | cut -c1-150
echo "== CL=1 rings 3/1"; GR_NMF_NO_CLUSTER=1 GR_NMF_RING_A=3 GR_NMF_RING_B=1 timeout 120 python tools/bench_nmf.py --ranks 32 --iters 10 --paths tcgen05 2>&1 | cut -c1-150
echo "== CL=2 default rings"; timeout 120 python tools/bench_nmf.py --ranks 4,16 --iters 10 --paths tcgen05 2>&1 | cut -c1-150
echo "== CL=2 rings 3/1"; GR_NMF_RING_A=3 GR_NMF_RING_B=1 timeout 120 python tools/bench_nmf.py --ranks 32 --iters 10 --paths tcgen05 2>&1 | cut -c1-150
echo "== CL=2 rings 2/1"; GR_NMF_RING_A=2 GR_NMF_RING_B=1 timeout 120 python tools/bench_nmf.py --ranks 32 --iters 10 --paths tcgen05 2>&1 | cut -c1-150
GR_NMF_NO_CLUSTER=1 GR_NMF_TRACE=1 GR_NMF_TRACE_FIRST=20 timeout 120 python tools/bench_nmf.py --n 2000000 --ranks 32 --iters 1 --paths tcgen05 > gpurun_out/nmf_single_trace_e16.log 2>&1
grep "^blk" gpurun_out/nmf_single_trace_e16.log | tail -8 | cut -c1-160
