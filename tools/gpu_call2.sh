#!/bin/bash
# 2-GPU validation: new single-GPU tests, sharded parity (peer + nccl), N=2 bench in both exchange forms
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m pytest tests/test_refex_gpu.py -m gpu -x -q -k "shard or broadcast or barrier or two_gpu" > gpurun_out/pytest_shard.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_shard.log
tail -5 gpurun_out/pytest_shard.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29701 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_peer.json 2> gpurun_out/bench_n2_peer.err; echo "bench peer rc=$?"
tail -c 1500 gpurun_out/bench_n2_peer.json; tail -5 gpurun_out/bench_n2_peer.err
GR_SHARD_EXCHANGE=nccl timeout 600 $TR --master-port 29702 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err; echo "bench nccl rc=$?"
tail -c 1500 gpurun_out/bench_n2_nccl.json; tail -5 gpurun_out/bench_n2_nccl.err
