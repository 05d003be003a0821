#!/bin/bash
# Tree kernel: timing of configs 3 / 4 / 4 + barrier and the tree parity tests.
TAG=${1:-r2t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python scripts/bench_humanoids.py > $OUT/humanoids.json 2> $OUT/humanoids.err
cat $OUT/humanoids.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('%-60s %.3f ms  %s kkt %s' % (d['config'][:60], d['ms_per_step'], d['status_counts'], d['kkt_selfcheck']['stationarity_over_scale_max']))"
tail -2 $OUT/humanoids.err
timeout 900 python -m pytest tests -m gpu -q -k "humanoid or tree or g1 or draco or golden" > $OUT/pytest_tree.log 2>&1; tail -3 $OUT/pytest_tree.log
