#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2c34_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c34_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c34_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c34_smoke.log
timeout 200 python tools/bench_encode.py > gpurun_out/r2c34_bench_encode.txt 2>&1; grep -o '"bins": [0-9]*, "ms": [0-9.]*' gpurun_out/r2c34_bench_encode.txt | tr '\n' ' '; echo
( time timeout 900 python bench.py > gpurun_out/r2c34_bench_n1.json 2> gpurun_out/r2c34_bench_n1.err ) 2> gpurun_out/r2c34_bench_n1.time; echo "bench rc=$?"; tail -3 gpurun_out/r2c34_bench_n1.time
python - <<'P'
import json
d = json.loads(open('gpurun_out/r2c34_bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'parity', d['parity']['ok'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
print('next', {k: (v.get('ms') or v.get('wall_s')) for k, v in d['next'].items()}, d['next']['extract_features_device_resident'].get('kernel_ms'))
for r in d['nmf']['per_rank']: print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items() if k in ('r','path','ms_per_iter','frac_of_hbm_peak','ms_per_iter_with_convergence_checks','ms_per_convergence_check')})
print(d['nmf'].get('parity'))
print(d['nmf']['rolx_epilogue'].get('model_selection_grid'))
print(d['nmf']['rolx_epilogue'].get('encode_node_role_factor'))
P
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c34_bench_ref.json 2> gpurun_out/r2c34_bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r2c34_bench_ref.json
