#!/bin/bash
# Build a variant of the library whose nmf_mu_tc.cu is compiled with extra -D flags (development:
# compile-time experiments on the NMF pair kernel).  usage: tools/build_variant.sh <name> <flags...>
# -> tools/_variants/lib_<name>.so, timed with GR_EXP_LIB=<that path> python tools/bench_nmf.py
set -e
cd "$(dirname "$0")/../graphrole_b200/csrc"
name="$1"; shift
mkdir -p ../../tools/_variants
make -s libgraphrole_b200.so
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 "$@" \
     -c -o /tmp/nmf_mu_tc_$name.o nmf_mu_tc.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/_variants/lib_$name.so \
     capi.o refex_aggregate.o peer.o nmf_mu.o /tmp/nmf_mu_tc_$name.o prune.o level0.o rolx_epilogue.o
echo "built tools/_variants/lib_$name.so ($*)"
