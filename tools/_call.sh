#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/exp_nvls_probe 320 > gpurun_out/r2c30_nvls_probe_n8.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r2c30_nvls_probe_n8.txt | tail -30
