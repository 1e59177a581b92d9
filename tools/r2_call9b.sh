#!/bin/bash
# round 2, call 9b (2 GPUs): final C3 bench at 2 GPUs (NUMA-bound staging), memcheck of a 2-rank run
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2c9b_topo.txt 2>&1
timeout 400 $TR --nproc-per-node 2 --master-port 29732 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c9_bench_n2.json 2> gpurun_out/r2c9_bench_n2.err; echo "bench n2 rc=$?"; tail -2 gpurun_out/r2c9_bench_n2.err | cut -c1-300
python - <<'PY'
import json
for line in open('gpurun_out/r2c9_bench_n2.json'):
    if line.startswith('{'):
        d=json.loads(line)
        pr=d['sharding'].get('per_rank',{})
        print(round(d['ms_per_step'],2), 'kern', pr.get('kernel_ms_per_level'), 'wait', pr.get('barrier_wait_ms_per_level'), d['sharding']['parallelism'][:90])
        print('   parity', {k: d['parity'].get(k) for k in ('max_rel_err','bit_identical_to_single_gpu','ok','error')}, 'e2e', d['e2e'])
PY
timeout 900 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 $TR --nproc-per-node 2 --master-port 29733 tools/check_sharded.py --size 30000 --depth 2 > gpurun_out/r2c9b_sanitizer_memcheck_sharded_n2.log 2>&1; echo "memcheck sharded rc=$?"; grep -E "ERROR SUMMARY|SHARDED|Invalid|MISMATCH" gpurun_out/r2c9b_sanitizer_memcheck_sharded_n2.log | tail -8
