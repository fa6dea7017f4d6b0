#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list + full captures of the top kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [quick]
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if [ "$2" != "quick" ]; then
  ( time timeout 1200 python -m pytest tests -m gpu -q ) > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1
fi
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 400 python bench.py --impl reference --steps 12 --warmup 5 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
# launch list of the same command (cold-cache, serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --ncu-short > $OUT/${TAG}_ncu_b.log 2>&1
# full capture of the k-NN pair and the residual kernel of two steady-state steps
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_knn_stencil|k_residual|k_knn' -s 60 -c 12 \
    -o $OUT/${TAG}_full -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --ncu-short > $OUT/${TAG}_ncu_f.log 2>&1
# dram bytes of one k-NN search pass (stencil + exact kernel of the first search pass captured) -> profiles/knn_traffic.json
python tools/ncu_summary.py traffic $OUT/${TAG}_full.ncu-rep $OUT/${TAG}_knn_traffic.json > /dev/null 2>&1
if [ -f better_fastlio2_b200/libfastlio_b200_trace.so ]; then
  FLB_LIB=better_fastlio2_b200/libfastlio_b200_trace.so timeout 300 python tools/trace_step.py --steps 20 --out $OUT/${TAG}_trace.json > $OUT/${TAG}_trace.txt 2>&1
fi
tail -3 $OUT/${TAG}_pytest.log 2>/dev/null; tail -2 $OUT/${TAG}_smoke.log 2>/dev/null; head -c 1200 $OUT/${TAG}_bench.json; echo; head -c 600 $OUT/${TAG}_bench_ref.json
