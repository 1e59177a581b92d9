#!/bin/bash
# 8-GPU box: sharded parity at 2 and 8 ranks, C3 bench at 8 and 2 ranks (fused exchange)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 2 --master-port 29721 tools/check_sharded.py --size 200000 --depth 3 > gpurun_out/check_sharded_n2.log 2>&1; echo "check n2 rc=$?"
grep "rank\|SHARDED" gpurun_out/check_sharded_n2.log | tail -5 | cut -c1-200
timeout 200 $TR --nproc-per-node 8 --master-port 29722 tools/check_sharded.py --size 200000 --depth 3 > gpurun_out/check_sharded_n8.log 2>&1; echo "check n8 rc=$?"
grep "SHARDED\|MISMATCH\|Error" gpurun_out/check_sharded_n8.log | tail -5 | cut -c1-200
for N in 8 2; do
  timeout 400 $TR --nproc-per-node $N --master-port 2974$N bench.py --gpus $N --steps 5 --warmup 3 2> gpurun_out/bench_n${N}_peer.err > gpurun_out/bench_n${N}_peer.json; echo "bench N=$N rc=$?"
  python -c "
import json
for line in open('gpurun_out/bench_n${N}_peer.json'):
    if line.startswith('{'):
        d=json.loads(line); print($N, d['ms_per_step'], d['value'], d['roofline']['kernel_ms_avg'], d['roofline']['frac'], d['config']['exchange'][:40], d['clocks'])"
  tail -2 gpurun_out/bench_n${N}_peer.err | cut -c1-200
done
