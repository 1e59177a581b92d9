#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2 8 11; do
  GR_NMF_TC_DEBUG=$dbg timeout 200 python tools/bench_nmf.py --ranks 8,32 --paths tcgen05 --iters 20 > gpurun_out/r2c16_nmf_dbg$dbg.txt 2>&1
  echo "debug=$dbg: $(grep -o '"r": [0-9]*, "iters": [0-9]*, "ms_per_iter": [0-9.]*' gpurun_out/r2c16_nmf_dbg$dbg.txt | tr '\n' ' ')"
done
for ring in "2 2" "3 1"; do
  set -- $ring
  GR_NMF_RING_A=$1 GR_NMF_RING_B=$2 timeout 200 python tools/bench_nmf.py --ranks 8,32 --paths tcgen05 --iters 20 > gpurun_out/r2c16_nmf_ring$1$2.txt 2>&1
  echo "rings $1/$2: $(grep -o '"r": [0-9]*, "iters": [0-9]*, "ms_per_iter": [0-9.]*' gpurun_out/r2c16_nmf_ring$1$2.txt | tr '\n' ' ')"
done
GR_NMF_RING_A=3 GR_NMF_RING_B=1 GR_NMF_TC_DEBUG=11 timeout 200 python tools/bench_nmf.py --ranks 8 --paths tcgen05 --iters 20 > gpurun_out/r2c16_nmf_ring31_dbg11.txt 2>&1
echo "rings 3/1 debug=11: $(grep -o '"r": [0-9]*, "iters": [0-9]*, "ms_per_iter": [0-9.]*' gpurun_out/r2c16_nmf_ring31_dbg11.txt | tr '\n' ' ')"
