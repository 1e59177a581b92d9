#!/bin/bash
# round 2, call 4 (8 GPUs): parity of every sharding mode at 8 ranks; C3 bench at 8 and 4 GPUs,
# node ranges only vs 2 column groups x node ranges
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29721 tools/check_sharded.py --size 200000 --depth 3 > gpurun_out/r2c4_check_n8.log 2>&1; echo "check n8 rc=$?"
grep "rank 0\|SHARDED\|MISMATCH\|Error" gpurun_out/r2c4_check_n8.log | tail -12 | cut -c1-300
run() {  # name nproc extra-env extra-args
  env $3 timeout 400 $TR --nproc-per-node $2 --master-port 297$((RANDOM % 90 + 10)) bench.py --gpus $2 --steps 5 --warmup 3 $4 > gpurun_out/r2c4_$1.json 2> gpurun_out/r2c4_$1.err; echo "$1 rc=$?"; tail -2 gpurun_out/r2c4_$1.err | cut -c1-300
}
run bench_n8 8 "X=1" ""
run bench_n8_c1 8 "GR_SHARD_COL_GROUPS=1" "--no-parity --no-e2e"
run bench_n4 4 "X=1" ""
run bench_n4_c2 4 "GR_SHARD_COL_GROUPS=2" "--no-parity --no-e2e"
python - <<'PY'
import json
for f in ('bench_n8','bench_n8_c1','bench_n4','bench_n4_c2'):
    for line in open('gpurun_out/r2c4_%s.json' % f):
        if line.startswith('{'):
            d=json.loads(line)
            pr=d['sharding'].get('per_rank',{})
            print(f, round(d['ms_per_step'],2), 'kern', pr.get('kernel_ms_per_level'), 'wait', pr.get('barrier_wait_ms_per_level'), 'C', d['sharding']['column_groups'], d['sharding']['parallelism'][:80])
            print('   parity', d.get('parity'), 'e2e', d.get('e2e'))
            print('   hist', [(h['kernel_ms']) for h in d['sharding'].get('balancing_history',[])])
PY
