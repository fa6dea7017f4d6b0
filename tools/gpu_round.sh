#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list + full captures of the top kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/${TAG}_gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1
timeout 600 python bench.py --steps 100 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_knn_stencil|k_residual|k_knn' -s 12 -c 6 \
    -o $OUT/${TAG}_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_f.log 2>&1
if [ -f better_fastlio2_b200/libfastlio_b200_trace.so ]; then
  FLB_LIB=better_fastlio2_b200/libfastlio_b200_trace.so timeout 300 python tools/trace_step.py --steps 20 --out $OUT/${TAG}_trace.json > $OUT/${TAG}_trace.txt 2>&1
fi
tail -3 $OUT/${TAG}_pytest.log; cat $OUT/${TAG}_smoke.log | tail -2; cat $OUT/${TAG}_bench.json
