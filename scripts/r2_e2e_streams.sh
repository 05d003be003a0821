#!/bin/bash
# e2e with one and two caller streams, host schedules 0 and 2 (library: two staging sets).
OUT=gpurun_out/${1:-r2z}
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "host" 2>&1 | tail -2
for NS in 1 2; do
  for MODE in 0 2; do
    PK_HOST_MODE=$MODE timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --e2e-streams $NS > $OUT/b_$NS_$MODE.json 2>/dev/null
    python -c "
import json; d=json.load(open('$OUT/b_$NS_$MODE.json')); print('streams $NS mode $MODE: e2e %.1f us/step (%.3e) regions %s bitwise %s' % (d['e2e']['ms_per_step']*1e3, d['e2e']['value'], [round(x/20*1e3) for x in d['e2e']['region_ms']], d['e2e']['bitwise_equal_to_device_path']))"
  done
done
