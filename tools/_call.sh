#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests -q -m gpu -k "two_gpu" > gpurun_out/r2c35_pytest_two_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c35_pytest_two_gpu.log
timeout 600 $TR --master-port 29761 bench.py --gpus 2 > gpurun_out/r2c35_bench_n2.json 2> gpurun_out/r2c35_bench_n2.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    d = json.loads(open('gpurun_out/r2c35_bench_n2.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'parity', d.get('parity', {}).get('ok'), d.get('parity', {}).get('bit_identical_to_single_gpu'), 'e2e', d.get('e2e', {}).get('ms_per_step'))
    print(json.dumps(d.get('nmf_row_sharded'))[:1200])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2c35_bench_n2.err').read()[-1500:])
P
timeout 300 $TR --master-port 29771 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2c35_bench_ref_n2.json 2> gpurun_out/r2c35_bench_ref_n2.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r2c35_bench_ref_n2.json
