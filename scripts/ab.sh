#!/bin/bash
# A/B timing of kernel variants selected by environment variables.
OUT=gpurun_out/${1:-ab}
mkdir -p $OUT
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 4000 --warmup 20 --no-cpu > $OUT/$name.json 2> $OUT/$name.err; python - <<PY
import json
try:
    d=json.load(open("$OUT/$name.json")); print("$name", "kernel_us %.2f"%(1e3*d["roofline"]["kernel_ms"]), "value %.3e"%d["value"], "e2e_us %.1f"%(1e3*d["e2e"]["ms_per_step"]), d["nonzero_status"])
except Exception as e: print("$name", "ERR", e)
PY
}
run plain PK_CHAIN_MODE=0
run compact_rps1 PK_CHAIN_MODE=1 PK_ROUNDS_PER_SYNC=1
run compact_rps2 PK_CHAIN_MODE=1 PK_ROUNDS_PER_SYNC=2
run compact_rps3 PK_CHAIN_MODE=1 PK_ROUNDS_PER_SYNC=3
run compact_rps8 PK_CHAIN_MODE=1 PK_ROUNDS_PER_SYNC=8
run chunk8k PK_HOST_CHUNK=8192
run chunk32k PK_HOST_CHUNK=32768
run chunk64k PK_HOST_CHUNK=65536
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5
PK_CHAIN_MODE=0 timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
