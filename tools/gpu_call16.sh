#!/bin/bash
# 4-GPU check: sharded parity tests and the C3 bench in both exchange forms
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_refex_gpu.py -m gpu -x -q -k "two_gpu or shard or barrier" 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29711 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_n4_peer.json 2> gpurun_out/bench_n4_peer.err; echo "bench peer rc=$?"
tail -c 1800 gpurun_out/bench_n4_peer.json; tail -3 gpurun_out/bench_n4_peer.err
GR_SHARD_EXCHANGE=nccl timeout 600 $TR --master-port 29712 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_n4_nccl.json 2> gpurun_out/bench_n4_nccl.err; echo "bench nccl rc=$?"
tail -c 600 gpurun_out/bench_n4_nccl.json; tail -3 gpurun_out/bench_n4_nccl.err
