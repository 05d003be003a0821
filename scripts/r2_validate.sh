#!/bin/bash
# Final single-GPU validation: the driver's three steps (gpu tests, smoke, bench at its settings) + reference arm.
TAG=${1:-r2w}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_20.json 2> $OUT/bench_20.err; echo "bench exit $?"; tail -2 $OUT/bench_20.err
python - <<PY
import json
d = json.load(open("$OUT/bench_20.json"))
r = json.load(open("$OUT/bench_reference.json"))
print("value %.3e (%.2f us/step) frac %.4f | two-in-flight %.2f us | e2e %.3e (%.1f us/step) probe %s | ref arm %.3e (threads %s quota %s) | e2e ratio %.1f" % (
    d["value"], d["ms_per_step"]*1e3, d["roofline"]["frac"], (d["roofline"]["two_batches_in_flight_ms_per_step"] or 0)*1e3,
    d["e2e"]["value"], d["e2e"]["ms_per_step"]*1e3, d["e2e"].get("schedule_probe_us_per_call"), r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["cfs_quota_cpus"], d["e2e"]["value"]/r["value"]))
print("cpu_baseline in-line:", d["cpu_baseline"]["value"], d["cpu_baseline"]["parallel_efficiency"])
for c in d["configs"]:
    print("  %-70s %.3f ms kkt %.2e" % (c["config"][:70], c["ms_per_step"], c["kkt_selfcheck"]["stationarity_over_scale_max"]))
PY
