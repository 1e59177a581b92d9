#!/bin/bash
# NMF tcgen05 kernel: where does the time go?  C5 (10M x 512, r = 32) with parts switched off,
# then the per-block timeline of CTA 0.
mkdir -p gpurun_out
for dbg in 0 1 2 4 8 12 13; do
  echo "== GR_NMF_TC_DEBUG=$dbg" >> gpurun_out/nmf_parts.log
  GR_NMF_TC_DEBUG=$dbg timeout 300 python tools/bench_nmf.py --ranks 32 --iters 10 --paths tcgen05 >> gpurun_out/nmf_parts.log 2>&1
done
cat gpurun_out/nmf_parts.log | cut -c1-200
GR_NMF_TRACE=1 GR_NMF_TRACE_FIRST=20 timeout 300 python tools/bench_nmf.py --n 2000000 --ranks 32 --iters 2 --paths tcgen05 > gpurun_out/nmf_trace.log 2>&1
tail -60 gpurun_out/nmf_trace.log | cut -c1-160
