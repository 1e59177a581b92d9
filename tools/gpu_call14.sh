#!/bin/bash
# round-1 final validation: whole GPU suite, C3 bench line, launch lists, ncu capture of the NMF kernel
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$? t=$(( $(date +%s) - t0 ))s"
tail -c 3000 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'nmf' -c 60 --csv \
    --log-file gpurun_out/launches_nmf.csv python tools/bench_nmf.py --ranks 32 --iters 4 --paths tcgen05 > gpurun_out/launches_nmf_run.log 2>&1
echo "ncu list rc=$? t=$(( $(date +%s) - t0 ))s"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nmf_fused_tc -s 3 -c 1 \
    -o gpurun_out/prof_nmf_tc python tools/bench_nmf.py --n 4000000 --ranks 32 --iters 2 --paths tcgen05 > gpurun_out/prof_nmf_run.log 2>&1
echo "ncu full rc=$? t=$(( $(date +%s) - t0 ))s"
ncu -i gpurun_out/prof_nmf_tc.ncu-rep --page raw --csv > gpurun_out/ncu_nmf_tc_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
