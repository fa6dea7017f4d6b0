#!/bin/bash
# Bench-only A/B (no tests): bash tools/gpu_ab_quick.sh <tag> ENV=a ENV=b ...   (each run: bench.py --steps 200, no CPU legs)
TAG=${1:-abq}; shift
OUT=gpurun_out; mkdir -p $OUT
for kv in "$@"; do
  name=$(echo "$kv" | tr '= /.' '____')
  env $kv timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  echo "$kv: $(python -c "import json; d=json.load(open('$OUT/${TAG}_${name}.json')); print(round(d['value'],1), round(d['e2e']['value'],1), d['kernel_ms_per_step']['knn'])")"
done
