#!/bin/bash
# Evidence set of a round in ONE GPU-box visit: tools/gpu_round.sh (bench both arms, ncu launch list + full capture, k-NN
# traffic, device timeline) + the other BASELINE configs at full size + a long-block consistency check.
# Usage: bash tools/gpu_final.sh <tag> [quick]     (quick: without the parity tests and smoke())
TAG=${1:-rz}
OUT=gpurun_out
bash tools/gpu_round.sh $TAG $2
timeout 400 python bench.py --config cfg3 > $OUT/${TAG}_bench_cfg3.json 2> $OUT/${TAG}_bench_cfg3.err
timeout 500 python bench.py --config cfg4 > $OUT/${TAG}_bench_cfg4.json 2> $OUT/${TAG}_bench_cfg4.err
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_steps200.json 2> $OUT/${TAG}_bench_steps200.err
head -c 400 $OUT/${TAG}_bench_cfg3.json; echo; head -c 400 $OUT/${TAG}_bench_cfg4.json; echo; head -c 300 $OUT/${TAG}_bench_steps200.json
