#!/bin/sh
# Build the DEBUG library with the device-side timeline (csrc/trace.cuh).  Not part of build(); never shipped as the
# product library.  Usage: tools/trace_build.sh  ->  better_fastlio2_b200/libfastlio_b200_trace.so
set -e
cd "$(dirname "$0")/.."
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -DFLB_TRACE \
  -Xcompiler -fPIC -shared -ccbin /usr/bin/g++ -o better_fastlio2_b200/libfastlio_b200_trace.so \
  better_fastlio2_b200/csrc/fastlio_b200.cu
