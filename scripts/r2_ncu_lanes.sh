#!/bin/bash
# ncu --set full capture of the chain kernel for the given PK_CHAIN_LANES values, summarised ON
# the GPU box (the reports are ~28 MB each, more than gpurun carries back): per variant one
# text file with the launch metrics, the stall reasons, the per-source-line attribution and
# the per-phase instruction counts.
# Usage: bash scripts/r2_ncu_lanes.sh tag "0 1 2"
TAG=${1:-r2c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for LANES in $2; do
  REP=/tmp/prof_lanes$LANES
  PK_CHAIN_LANES=$LANES timeout 600 ncu --set full --clock-control none -k regex:"ik_chain|ik_coop" -s 6 -c 1 \
      -o $REP python bench.py --steps 8 --warmup 3 --regions 1 --no-cpu --no-configs --nbuf 4 > $OUT/ncu_lanes$LANES.log 2>&1
  if [ "$LANES" = "0" ]; then PAT=ik_chain_kernelILi6ELi1E; else PAT=ik_coop_kernelILi6ELi1ELi${LANES}E; fi
  python scripts/profile_summary.py $REP.ncu-rep pink_b200/libpink_b200.so $PAT > $OUT/summary_lanes$LANES.txt 2>&1
  if [ "$LANES" != "0" ]; then
    python scripts/ncu_phases.py pink_b200/libpink_b200.so $PAT $REP.ncu-rep $((2048 * LANES)) >> $OUT/summary_lanes$LANES.txt 2>&1
  fi
  # dram traffic of the launch for profiles/traffic.json
  ncu -i $REP.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin)); h = rows[0]; r = rows[2]
rd, wr = h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum')
u = rows[1]
print('dram_read', r[rd], u[rd], 'dram_write', r[wr], u[wr])" >> $OUT/summary_lanes$LANES.txt 2>&1
  rm -f $REP.ncu-rep
  head -12 $OUT/summary_lanes$LANES.txt | tail -9
done
