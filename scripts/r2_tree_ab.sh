#!/bin/bash
# A/B of tree-kernel builds (resident CTAs per SM the register allocation aims at).
OUT=gpurun_out/${1:-r2x}
mkdir -p $OUT
cp pink_b200/libpink_b200.so /tmp/lib_orig.so
for V in mb8 mb10; do
  cp build/lib_$V.so pink_b200/libpink_b200.so
  echo "== $V"
  timeout 600 python scripts/bench_humanoids.py 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('%-60s %.3f ms' % (d['config'][:60], d['ms_per_step']))"
done
cp /tmp/lib_orig.so pink_b200/libpink_b200.so
