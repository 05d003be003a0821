#!/bin/bash
# ncu --set full capture of the tree kernel on config 4 (G1-class + self-collision barrier:
# the warp-cooperative dual QP).  bench_humanoids.py launches ik_tree_kernel 13 times per
# configuration; the capture takes one launch of the third configuration.
OUT=gpurun_out/${1:-ncu_tree_dual}
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ik_tree_kernel -s 30 -c 1 \
    -o $OUT/prof_tree_dual python scripts/bench_humanoids.py > $OUT/ncu.log 2>&1
cp pink_b200/libpink_b200.so $OUT/libpink_b200.so
tail -4 $OUT/ncu.log | cut -c1-200
