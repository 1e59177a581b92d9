#!/bin/bash
mkdir -p gpurun_out
GR_NMF_NARROW_P2=1 timeout 300 python tools/bench_nmf.py --ranks 4,8,16,32 --paths tcgen05 --iters 20 > gpurun_out/r2c18_nmf_narrow.txt 2>&1; echo "narrow: $(grep -o '"r": [0-9]*, "iters": [0-9]*, "ms_per_iter": [0-9.]*' gpurun_out/r2c18_nmf_narrow.txt | tr '\n' ' ')"
timeout 300 python tools/bench_nmf.py --ranks 4,8,16,32 --paths tcgen05 --iters 20 > gpurun_out/r2c18_nmf_wide.txt 2>&1; echo "wide: $(grep -o '"r": [0-9]*, "iters": [0-9]*, "ms_per_iter": [0-9.]*' gpurun_out/r2c18_nmf_wide.txt | tr '\n' ' ')"
for dbg in 1 2 11; do
  GR_NMF_TC_DEBUG=$dbg timeout 200 python tools/bench_nmf.py --ranks 8,32 --paths tcgen05 --iters 20 > gpurun_out/r2c18_nmf_dbg$dbg.txt 2>&1
  echo "debug=$dbg: $(grep -o '"r": [0-9]*, "iters": [0-9]*, "ms_per_iter": [0-9.]*' gpurun_out/r2c18_nmf_dbg$dbg.txt | tr '\n' ' ')"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nmf_fused_tc -s 2 -c 1 -o gpurun_out/r2c18_ncu_nmf_tc -f python tools/bench_nmf.py --n 4000000 --ranks 32 --paths tcgen05 --iters 3 > gpurun_out/r2c18_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2c18_ncu.log
