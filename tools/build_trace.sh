#!/bin/bash
# Instrumented tuning build: one translation unit (default vae_engine; `tools/build_trace.sh spade` for the generator's) compiled with
# -DSLN_TC_TRACE (per-phase SM-clock stamps of CTA (0,0,0), read by sln_debug_tc_trace[_spade] / tools/bench_contract.py /
# tools/trace_spade.py) and linked with the regular objects -> sln_b200/libsln_b200_trace.so.
# Use with SLN_LIB_PATH=sln_b200/libsln_b200_trace.so.  Never the shipped library.
set -e
cd "$(dirname "$0")/.."
TU=${1:-vae_engine}
mkdir -p sln_b200/build_trace
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DSLN_TC_TRACE -I include -c -o sln_b200/build_trace/$TU.o sln_b200/csrc/$TU.cu
objs=""
for f in vae_engine runtime raster spade collate refine_loss scene; do
  if [ $f != $TU ]; then objs="$objs sln_b200/build/$f.o"; fi
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o sln_b200/libsln_b200_trace.so sln_b200/build_trace/$TU.o $objs
echo built sln_b200/libsln_b200_trace.so
