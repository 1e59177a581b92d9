#!/bin/bash
# round 2, call 9a (8 GPUs): parity of every sharding mode at 8 ranks, final C3 bench at 8 and 4 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29721 tools/check_sharded.py --size 200000 --depth 3 > gpurun_out/r2c9_check_n8.log 2>&1; echo "check n8 rc=$?"
grep "SHARDED\|MISMATCH\|Error" gpurun_out/r2c9_check_n8.log | tail -6 | cut -c1-300
for N in 8 4; do
  timeout 400 $TR --nproc-per-node $N --master-port 2973$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2c9_bench_n$N.json 2> gpurun_out/r2c9_bench_n$N.err; echo "bench n$N rc=$?"; tail -2 gpurun_out/r2c9_bench_n$N.err | cut -c1-300
done
python - <<'PY'
import json
for f in ('bench_n8','bench_n4'):
    for line in open('gpurun_out/r2c9_%s.json' % f):
        if line.startswith('{'):
            d=json.loads(line)
            pr=d['sharding'].get('per_rank',{})
            print(f, round(d['ms_per_step'],2), 'kern', pr.get('kernel_ms_per_level'), 'wait', pr.get('barrier_wait_ms_per_level'), d['sharding']['parallelism'][:90])
            print('   parity', {k: d['parity'].get(k) for k in ('max_rel_err','bit_identical_to_single_gpu','ok','error')}, 'e2e', {k: d['e2e'].get(k) for k in ('ms_per_step','h2d_bytes_per_step','d2h_bytes_per_step','matches_device_path','error')})
            print('   hist', [(h['kernel_ms']) for h in d['sharding'].get('balancing_history',[])])
PY
