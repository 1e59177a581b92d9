#!/bin/bash
# round 2, call 11 (1 GPU): what the driver runs -- GPU suite, smoke, bench (both arms)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2c11_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 gpurun_out/r2c11_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c11_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2c11_smoke.log
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c11_bench_n1.json 2> gpurun_out/r2c11_bench_n1.err ) 2> gpurun_out/r2c11_bench_n1.time; echo "bench rc=$?"; tail -3 gpurun_out/r2c11_bench_n1.err; grep real gpurun_out/r2c11_bench_n1.time
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2c11_bench_ref.json 2> gpurun_out/r2c11_bench_ref.err ) 2> gpurun_out/r2c11_bench_ref.time; echo "ref arm rc=$?"; tail -3 gpurun_out/r2c11_bench_ref.err; grep real gpurun_out/r2c11_bench_ref.time; cut -c1-400 gpurun_out/r2c11_bench_ref.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c11_bench_n1.json').read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'parity', d['parity']['ok'], d['parity']['max_rel_err'])
print('e2e', d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['kind'], d['cpu_baseline']['value'])
print('next', json.dumps(d['next'])[:1800])
print('c2', d['c2']['ms_per_step'], d['c2']['roofline']['frac'], d['c2']['parity']['ok'])
for row in d['nmf']['per_rank']: print(row)
print(json.dumps(d['nmf']['rolx_epilogue'])[:900])
PY
