#!/bin/bash
# Instrumented tuning build: vae_engine.cu with -DSLN_TC_TRACE (per-phase SM-clock stamps of CTA (0,0,0), read by
# sln_debug_tc_trace / tools/bench_contract.py) linked with the regular objects -> sln_b200/libsln_b200_trace.so.
# Use with SLN_LIB_PATH=sln_b200/libsln_b200_trace.so.  Never the shipped library.
set -e
cd "$(dirname "$0")/.."
mkdir -p sln_b200/build_trace
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DSLN_TC_TRACE -I include -c -o sln_b200/build_trace/vae_engine.o sln_b200/csrc/vae_engine.cu
objs=""
for f in runtime raster spade collate refine_loss scene; do objs="$objs sln_b200/build/$f.o"; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o sln_b200/libsln_b200_trace.so sln_b200/build_trace/vae_engine.o $objs
echo built sln_b200/libsln_b200_trace.so
