#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2c42_level0_variants.txt
for cfg in 8 8p 4 4p 16 16p; do
  GR_LEVEL0_TRI=$cfg timeout 200 python tools/exp_level0_variants.py 2>&1 | tail -2 >> gpurun_out/r2c42_level0_variants.txt
done
cat gpurun_out/r2c42_level0_variants.txt
