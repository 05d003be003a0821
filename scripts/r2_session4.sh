#!/bin/bash
# Full GPU test-suite on the current build, bench at the driver's settings for the default
# chain kernel and the L = 1 sub-warp kernel, ncu captures of the lane variants.
TAG=${1:-r2k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_20.json 2> $OUT/bench_20.err; echo "bench exit $?"
for LANES in 1; do
  PK_CHAIN_LANES=$LANES timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs > $OUT/bench_lanes$LANES.json 2> $OUT/bench_lanes$LANES.err
done
python - <<PY
import json
for f in ("bench_20", "bench_lanes1"):
    try:
        d = json.load(open("$OUT/%s.json" % f))
        print(f, "%.2f us/step value %.3e e2e %.1f us" % (d["ms_per_step"]*1e3, d["value"], d["e2e"]["ms_per_step"]*1e3), d.get("cpu_baseline", {}).get("value"), [ (c["ms_per_step"]) for c in d.get("configs", []) if isinstance(c, dict)])
    except Exception as e:
        print(f, "failed", e)
PY
bash scripts/r2_ncu_lanes.sh $TAG "0 1 4 8" > $OUT/ncu.log 2>&1
ls -la $OUT | head -30
