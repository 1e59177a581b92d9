#!/bin/bash
# 4-GPU: sharded parity (2 and 4 ranks, full output), C3 bench with the default and a lower hub threshold
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29721 tools/check_sharded.py --n 200000 --levels 3 > gpurun_out/check_sharded_n2.log 2>&1; echo "check n2 rc=$?"
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/check_sharded_n2.log | tail -12 | cut -c1-300
timeout 300 $TR --nproc-per-node 4 --master-port 29722 tools/check_sharded.py --n 200000 --levels 3 > gpurun_out/check_sharded_n4.log 2>&1; echo "check n4 rc=$?"
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/check_sharded_n4.log | tail -8 | cut -c1-300
for cfg in "2048 1024" "256 256"; do
  set -- $cfg
  echo "== N=4 peer, hub threshold $1 segment $2"
  GR_REFEX_HUB_THRESHOLD=$1 GR_REFEX_HUB_SEGMENT=$2 timeout 600 $TR --nproc-per-node 4 --master-port 2973$((RANDOM % 10)) bench.py --gpus 4 --steps 5 --warmup 3 2>/dev/null > gpurun_out/bench_n4_peer_$1.json
  python -c "
import json
for line in open('gpurun_out/bench_n4_peer_$1.json'):
    if line.startswith('{'):
        d=json.loads(line); print(d['ms_per_step'], d['roofline']['kernel_ms_avg'], d['roofline']['frac'], d['config']['exchange'][:20])"
done
