#!/bin/bash
# Lanes-per-instance study of the chain kernel: parity tests + device-resident timing for
# PK_CHAIN_LANES = 0 (round-1 kernel), 1, 2, 4, 8, and one ncu --set full capture per variant.
# Usage: bash scripts/r2_lanes.sh [tag] [ncu: 0|1]
TAG=${1:-r2b}
NCU=${2:-1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for LANES in 0 1 2 4 8; do
  export PK_CHAIN_LANES=$LANES
  if [ $LANES -le 2 ]; then
    timeout 600 python -m pytest tests -m gpu -q -x -k "ur5 or chain or rollout or golden or scenarios or hostile or examples" > $OUT/pytest_lanes$LANES.log 2>&1
    echo "lanes $LANES pytest exit $?" >> $OUT/pytest_lanes$LANES.log
  else
    timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ur5" > $OUT/pytest_lanes$LANES.log 2>&1
    echo "lanes $LANES pytest exit $?" >> $OUT/pytest_lanes$LANES.log
  fi
  timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu --no-configs > $OUT/bench_lanes$LANES.json 2> $OUT/bench_lanes$LANES.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_lanes$LANES.json"))
    print("lanes $LANES: %.2f us/step, eager %.2f us, e2e %.1f us, nonzero status %d" % (d["ms_per_step"]*1e3, d["roofline"]["eager_ms_per_step"]*1e3, d["e2e"]["ms_per_step"]*1e3, d["nonzero_status"]))
except Exception as e:
    print("lanes $LANES: bench failed", e)
PY
  tail -2 $OUT/pytest_lanes$LANES.log
done
if [ "$NCU" = "1" ]; then
  for LANES in 0 1 2 4 8; do
    PK_CHAIN_LANES=$LANES timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ik_chain|ik_coop" -s 6 -c 1 \
        -o $OUT/prof_lanes$LANES python bench.py --steps 8 --warmup 3 --regions 1 --no-cpu --no-configs --nbuf 4 > $OUT/ncu_lanes$LANES.log 2>&1
  done
fi
