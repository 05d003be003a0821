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
run sps1 PK_STEPS_PER_SYNC=1
run sps2 PK_STEPS_PER_SYNC=2
run sps4 PK_STEPS_PER_SYNC=4
run sps8 PK_STEPS_PER_SYNC=8
run sps32 PK_STEPS_PER_SYNC=32
run chunk16k PK_HOST_CHUNK=16384
run chunk32k PK_HOST_CHUNK=32768
python scripts/pcie_probe.py > $OUT/pcie.txt 2>&1; cat $OUT/pcie.txt
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
