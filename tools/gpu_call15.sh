#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_prune_level0_gpu.py -m gpu -x -q > gpurun_out/pytest_next.log 2>&1; echo "pytest-next rc=$?"
tail -4 gpurun_out/pytest_next.log
timeout 600 python bench.py --steps 3 --no-e2e --no-cpu-baseline --no-nmf > gpurun_out/bench_next.json 2> gpurun_out/bench_next.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_next.json')); print(json.dumps(d['next'], indent=1))"
tail -3 gpurun_out/bench_next.err
