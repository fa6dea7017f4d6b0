#!/bin/bash
# Last check of a round on the final build: parity tests, smoke(), the default bench line (with roofline.traffic).
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-last}
( time timeout 900 python -m pytest tests -m gpu -q ) > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -3 $OUT/${TAG}_pytest.log; tail -2 $OUT/${TAG}_smoke.log; head -c 600 $OUT/${TAG}_bench.json
