#!/bin/bash
# round-1 re-validation: GPU tests, bench (C3), launch list, full ncu capture of the gather kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_c3.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'refex|hub' -c 40 --csv \
    --log-file gpurun_out/launches_c3.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-nmf > gpurun_out/launches_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:refex_gather -s 4 -c 2 \
    -o gpurun_out/prof_refex python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-nmf > gpurun_out/prof_run.log 2>&1
ls -la gpurun_out
