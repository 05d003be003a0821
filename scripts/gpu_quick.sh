#!/bin/bash
# Short GPU session: tests + humanoid timings (no ncu).  Usage: bash scripts/gpu_quick.sh [tag]
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 600 python scripts/bench_humanoids.py > $OUT/humanoids.json 2> $OUT/humanoids.err
timeout 300 python bench.py --steps 2000 --warmup 20 --no-cpu > $OUT/bench.json 2> $OUT/bench.err
tail -15 $OUT/pytest_gpu.log; cat $OUT/humanoids.json; tail -3 $OUT/humanoids.err; cat $OUT/bench.json
