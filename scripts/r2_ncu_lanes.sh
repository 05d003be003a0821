#!/bin/bash
# ncu --set full capture of the chain kernel for the given PK_CHAIN_LANES values.
# Usage: bash scripts/r2_ncu_lanes.sh tag "0 1 2"
TAG=${1:-r2c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for LANES in $2; do
  PK_CHAIN_LANES=$LANES timeout 600 ncu --set full --clock-control none -k regex:"ik_chain|ik_coop" -s 6 -c 1 \
      -o $OUT/prof_lanes$LANES python bench.py --steps 8 --warmup 3 --regions 1 --no-cpu --no-configs --nbuf 4 > $OUT/ncu_lanes$LANES.log 2>&1
  tail -2 $OUT/ncu_lanes$LANES.log
done
ls -la $OUT
