#!/bin/bash
# round 2, call 5 (1 GPU): NMF sender warps (parity + timing), U = 2 narrow gathers, ncu of the
# d = 32 gather, the whole default bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_nmf_gpu.py -x -q -k "tensor_core_path or fixed_iterations" > gpurun_out/r2c5_nmf_quick.log 2>&1; echo "nmf quick rc=$?"; tail -3 gpurun_out/r2c5_nmf_quick.log
timeout 900 python -m pytest tests/test_nmf_gpu.py -x -q -s > gpurun_out/r2c5_nmf.log 2>&1; echo "nmf rc=$?"; grep -E "pair kernel vs|passed|failed|Error" gpurun_out/r2c5_nmf.log | tail -12
timeout 300 python tools/bench_nmf.py --ranks 8,32 --paths tcgen05 --iters 20 > gpurun_out/r2c5_nmf_send.txt 2>&1; echo "bench_nmf send rc=$?"; grep ms_per_iter gpurun_out/r2c5_nmf_send.txt | cut -c1-200
GR_NMF_NO_SENDERS=1 timeout 300 python tools/bench_nmf.py --ranks 8,32 --paths tcgen05 --iters 20 > gpurun_out/r2c5_nmf_nosend.txt 2>&1; echo "bench_nmf nosend rc=$?"; grep ms_per_iter gpurun_out/r2c5_nmf_nosend.txt | cut -c1-200
timeout 300 python tools/exp_narrow_u.py > gpurun_out/r2c5_narrow_u.txt 2>&1; echo "narrow rc=$?"; cat gpurun_out/r2c5_narrow_u.txt | tail -8
timeout 300 ncu --set full --clock-control none --import-source on -k regex:refex_gather -c 1 -o gpurun_out/r2c5_ncu_gather_d32 -f python tools/exp_narrow_u.py --one --widths 32 > gpurun_out/r2c5_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2c5_ncu.log
timeout 300 python -m pytest tests/test_rolx_gpu.py tests/test_prune_level0_gpu.py -x -q > gpurun_out/r2c5_rolx_prune.log 2>&1; echo "rolx+prune rc=$?"; tail -3 gpurun_out/r2c5_rolx_prune.log
timeout 900 python bench.py --steps 10 > gpurun_out/r2c5_bench_n1.json 2> gpurun_out/r2c5_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c5_bench_n1.err; cut -c1-600 gpurun_out/r2c5_bench_n1.json
