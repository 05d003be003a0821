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
run mode0 PK_HOST_MODE=0
run mode2 PK_HOST_MODE=2
run mode2_c16 PK_HOST_MODE=2 PK_HOST_CHUNK=16384
run mode2_c64 PK_HOST_MODE=2 PK_HOST_CHUNK=65536
run mode0b PK_HOST_MODE=0
run mode2b PK_HOST_MODE=2
