#!/bin/bash
# Instrumented build of the library for tools/tc_timing.py (phase-cycle counters in the fused tensor-core kernel).
# Needs a normal build first (python __graft_entry__.py): only bfx_k_tc.cu is recompiled with -DBFX_TIMING.
set -e
cd "$(dirname "$0")/../bayesflux.jl_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DBFX_TIMING \
  -c -o /tmp/bfx_k_tc_timing.o bfx_k_tc.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/${1:-libbfx_timing.so} bfx_api.o bfx_kernels.o \
  bfx_k_small.o bfx_k_256.o /tmp/bfx_k_tc_timing.o bfx_k_rec.o bfx_stream.o bfx_stream_tc.o bfx_diag.o
echo built tools/${1:-libbfx_timing.so}
