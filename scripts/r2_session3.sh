#!/bin/bash
# ncu captures of the remaining lane variants, bench at the driver's settings (e2e warm-up fix),
# reference arm with the futex pool.
TAG=${1:-r2d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
bash scripts/r2_ncu_lanes.sh $TAG "0 4 8" > $OUT/ncu.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_20.json 2> $OUT/bench_20.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2>> $OUT/bench_20.err
OC_POOL_PIN=0 timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference_nopin.json 2>> $OUT/bench_20.err
python - <<PY
import json
d=json.load(open("$OUT/bench_20.json"))
print("value %.3e  %.2f us/step; e2e %.1f us/step warm %d regions %s" % (d["value"], d["ms_per_step"]*1e3, d["e2e"]["ms_per_step"]*1e3, d["e2e"]["warmup_calls"], [round(x/20*1e3) for x in d["e2e"]["region_ms"]]))
print("cpu_baseline", d.get("cpu_baseline"))
for f in ("bench_reference", "bench_reference_nopin"):
    r=json.load(open("$OUT/%s.json" % f)); print(f, r["value"], r["cpu_baseline"]["parallel_efficiency"], r["cpu_baseline"]["one_core"])
PY
tail -3 $OUT/bench_20.err
