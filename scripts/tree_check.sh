#!/bin/bash
OUT=gpurun_out/${1:-tree}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
timeout 600 python scripts/bench_humanoids.py > $OUT/humanoids_tree.json 2> $OUT/humanoids.err; cat $OUT/humanoids_tree.json; tail -3 $OUT/humanoids.err
PK_TREE=0 timeout 600 python scripts/bench_humanoids.py > $OUT/humanoids_generic.json 2>> $OUT/humanoids.err; cat $OUT/humanoids_generic.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ik_tree -s 3 -c 1 -o $OUT/prof_tree \
    python scripts/bench_humanoids.py > $OUT/ncu_tree.log 2>&1
cp pink_b200/libpink_b200.so $OUT/
