#!/bin/bash
# Device-side timeline of the graph-launched step (debug library) + one bench line.  Usage: bash tools/gpu_trace.sh <tag>
TAG=${1:-tr}
OUT=gpurun_out
mkdir -p $OUT
FLB_LIB=better_fastlio2_b200/libfastlio_b200_trace.so timeout 300 python tools/trace_step.py --steps 20 --out $OUT/${TAG}_trace.json > $OUT/${TAG}_trace.txt 2>&1
cat $OUT/${TAG}_trace.txt | tail -45
timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -3 $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json
