#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -q ) > $OUT/r3y_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/r3y_pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/r3y_smoke.log 2>&1
timeout 600 python bench.py > $OUT/r3y_bench.json 2> $OUT/r3y_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r3y_smoke_launches.csv python __graft_entry__.py smoke > $OUT/r3y_smoke_ncu.log 2>&1
tail -3 $OUT/r3y_pytest.log; tail -2 $OUT/r3y_smoke.log; head -c 600 $OUT/r3y_bench.json
