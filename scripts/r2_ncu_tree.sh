#!/bin/bash
# ncu --set full of the tree kernel on config 3 (Draco3-class), summarised on the box.
TAG=${1:-r2v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
REP=/tmp/prof_tree
timeout 600 ncu --set full --clock-control none -k regex:ik_tree_kernel -s 3 -c 1 -o $REP python scripts/bench_humanoids.py > $OUT/ncu_tree.log 2>&1
python scripts/profile_summary.py $REP.ncu-rep pink_b200/libpink_b200.so ik_tree_kernel > $OUT/summary_tree_config3.txt 2>&1
sed -i 's/"32"\]/"60"]/' /dev/null
rm -f $REP.ncu-rep
head -40 $OUT/summary_tree_config3.txt
