#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29751 tools/check_sharded_nmf.py --rows 300000 --bench > gpurun_out/r2c28_check_nmf_n8.txt 2>&1; echo "check8 rc=$?"; grep -E '^\{|SHARDED' gpurun_out/r2c28_check_nmf_n8.txt | cut -c1-330
timeout 400 $TR --nproc-per-node 8 --master-port 29761 bench.py --gpus 8 > gpurun_out/r2c28_bench_n8.json 2> gpurun_out/r2c28_bench_n8.err; echo "bench8 rc=$?"
python - <<'P'
import json
try:
    d = json.loads(open('gpurun_out/r2c28_bench_n8.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'parity', d.get('parity', {}).get('ok'), d.get('parity', {}).get('bit_identical_to_single_gpu'), 'e2e', d.get('e2e', {}).get('ms_per_step'))
    print(json.dumps(d.get('nmf_row_sharded'))[:900])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2c28_bench_n8.err').read()[-1500:])
P
timeout 200 $TR --nproc-per-node 2 --master-port 29771 tools/check_sharded_nmf.py --rows 300000 > gpurun_out/r2c28_check_nmf_n2.txt 2>&1; echo "check2 rc=$?"; grep -E '^\{|SHARDED' gpurun_out/r2c28_check_nmf_n2.txt | cut -c1-330
timeout 200 $TR --nproc-per-node 4 --master-port 29781 tools/check_sharded_nmf.py --rows 300000 --bench > gpurun_out/r2c28_check_nmf_n4.txt 2>&1; echo "check4 rc=$?"; grep -E 'bench|SHARDED' gpurun_out/r2c28_check_nmf_n4.txt | cut -c1-330
