#!/bin/bash
# A/B timing of kernel variants selected by environment variables.
OUT=gpurun_out/${1:-ab}
mkdir -p $OUT
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3200 --warmup 20 --no-cpu > $OUT/$name.json 2> $OUT/$name.err; python - <<PY
import json
try:
    d=json.load(open("$OUT/$name.json")); print("$name", "kernel_us %.2f"%(1e3*d["roofline"]["kernel_ms"]), "value %.3e"%d["value"], "eager_us %.1f"%(1e3*d["roofline"]["eager_ms_per_step"]), "e2e_us %.1f"%(1e3*d["e2e"]["ms_per_step"]), d["nonzero_status"])
except Exception as e: print("$name", "ERR", e); print(open("$OUT/$name.err").read()[-800:])
PY
}
run base
run probe PK_PROBE_SKIP_ROUNDS=1
run base2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -3
