#!/bin/bash
# round 2, call 3 (2 GPUs): sharded parity in every mode, bench at N=2 and N=1 (tiny first)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_rolx_gpu.py -x -q -k "select_model or roles_and" > gpurun_out/r2c3_rolx.log 2>&1; echo "rolx rc=$?"; tail -3 gpurun_out/r2c3_rolx.log
timeout 300 python bench.py --workload tiny --steps 3 --no-nmf > gpurun_out/r2c3_bench_tiny.json 2> gpurun_out/r2c3_bench_tiny.err; echo "bench tiny rc=$?"; tail -3 gpurun_out/r2c3_bench_tiny.err; cut -c1-1200 gpurun_out/r2c3_bench_tiny.json
timeout 400 $TR --nproc-per-node 2 --master-port 29721 tools/check_sharded.py --size 200000 --depth 3 > gpurun_out/r2c3_check_n2.log 2>&1; echo "check n2 rc=$?"
grep "rank 0\|SHARDED\|MISMATCH\|Error" gpurun_out/r2c3_check_n2.log | tail -12 | cut -c1-400
timeout 400 $TR --nproc-per-node 2 --master-port 29722 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c3_bench_n2.json 2> gpurun_out/r2c3_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r2c3_bench_n2.err | cut -c1-300
GR_SHARD_COL_GROUPS=2 timeout 400 $TR --nproc-per-node 2 --master-port 29723 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c3_bench_n2_c2.json 2> gpurun_out/r2c3_bench_n2_c2.err; echo "bench n2 C=2 rc=$?"; tail -3 gpurun_out/r2c3_bench_n2_c2.err | cut -c1-300
python - <<'PY'
import json
for f in ('r2c3_bench_n2.json','r2c3_bench_n2_c2.json'):
    for line in open('gpurun_out/'+f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, d['ms_per_step'], d['roofline']['kernel_ms_avg'], d['roofline']['frac'], d['sharding'].get('per_rank'), d.get('parity'), d.get('e2e'))
PY
