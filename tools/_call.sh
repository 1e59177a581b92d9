#!/bin/bash
# round-end validation on one B200: GPU test suite, smoke, the default bench, and an ncu capture
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2c40_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c40_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c40_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c40_smoke.log
( time timeout 900 python bench.py > gpurun_out/r2c40_bench_n1.json 2> gpurun_out/r2c40_bench_n1.err ) 2> gpurun_out/r2c40_bench_n1.time; echo "bench rc=$?"; tail -3 gpurun_out/r2c40_bench_n1.time
python - <<'P'
import json
d = json.loads(open('gpurun_out/r2c40_bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'parity', d['parity']['ok'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
print('next', {k: (v.get('ms') or v.get('wall_s')) for k, v in d['next'].items()}, d['next']['extract_features_device_resident'].get('kernel_ms'))
print('nmf parity', d['nmf'].get('parity', {}).get('ok'), 'roofline', {k: round(v['frac'], 3) for k, v in d['nmf'].get('roofline', {}).items()})
g = d['nmf']['rolx_epilogue']
print('grid', g.get('model_selection_grid'), 'warmup', g.get('linalg_warmup_s'))
P
timeout 300 ncu --set full --clock-control none --import-source on -k regex:triangle_kernel -c 1 -o gpurun_out/r2c40_ncu_triangle -f python - > gpurun_out/r2c40_ncu.log 2>&1 <<'P'
import sys, os
sys.path.insert(0, os.getcwd())
from graphrole_b200.graph.generators import barabasi_albert_csr
from graphrole_b200.graph import level0
g = barabasi_albert_csr(4_000_000, 20, seed=0, device='cuda:0')
level0.device_features(g)
P
echo "ncu rc=$?"
