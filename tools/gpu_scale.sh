#!/bin/bash
# Multi-GPU replicas (BASELINE cfg5 shape): one rank per GPU under torchrun, as the driver launches it.
# Usage (gpurun --gpus N): bash tools/gpu_scale.sh <tag> <N>
TAG=${1:-sc}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 50 --warmup 5 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
tail -3 $OUT/${TAG}_bench_n$N.err; cat $OUT/${TAG}_bench_n$N.json
