#!/bin/bash
# A/B of a tuning knob on the GPU box: parity tests first, then bench.py per setting.  Usage: bash tools/gpu_ab.sh <tag> ENV=a ENV=b ...
TAG=${1:-ab}; shift
OUT=gpurun_out
mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -15 $OUT/${TAG}_pytest.log
for kv in "$@"; do
  name=$(echo "$kv" | tr '= /.' '____')
  env $kv timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_${name}.json 2> $OUT/${TAG}_bench_${name}.err
  echo "$kv: $(python -c "import json,sys; d=json.load(open('$OUT/${TAG}_bench_${name}.json')); print(d['value'], d['e2e']['value'], d['kernel_ms_per_step'])")"
done
