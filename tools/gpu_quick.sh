#!/bin/bash
# Short GPU-box visit for an A/B: parity tests, one bench line, the device-side timeline.  Usage: bash tools/gpu_quick.sh <tag>
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -q -x ) > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
if [ -f better_fastlio2_b200/libfastlio_b200_trace.so ]; then
  FLB_LIB=better_fastlio2_b200/libfastlio_b200_trace.so timeout 300 python tools/trace_step.py --steps 20 --out $OUT/${TAG}_trace.json > $OUT/${TAG}_trace.txt 2>&1
fi
tail -4 $OUT/${TAG}_pytest.log; head -c 700 $OUT/${TAG}_bench.json; echo; tail -12 $OUT/${TAG}_trace.txt
