#!/bin/bash
mkdir -p gpurun_out
for cfg in "2048 1024" "512 256" "256 256" "128 128" "64 64"; do
  set -- $cfg
  echo "== hub threshold $1 segment $2"
  GR_REFEX_HUB_THRESHOLD=$1 GR_REFEX_HUB_SEGMENT=$2 timeout 300 python bench.py --steps 3 --no-e2e --no-cpu-baseline --no-nmf --no-next 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms_avg'], d['roofline']['frac'])"
done
GR_REFEX_HUB_THRESHOLD=128 GR_REFEX_HUB_SEGMENT=128 timeout 600 python -m pytest tests/test_refex_gpu.py -m gpu -x -q 2>&1 | tail -2
