#!/bin/bash
# One GPU session: tests, smoke, bench, ncu launch list + one full capture.
# Usage (from the repo root on the GPU box): bash scripts/gpu_check.sh [tag]
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
echo "smoke exit $?" >> $OUT/smoke.log
timeout 600 python bench.py --steps 20000 --warmup 50 > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?" >> $OUT/bench.err
timeout 600 python scripts/bench_humanoids.py > $OUT/humanoids.json 2>> $OUT/bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_reference.json 2>> $OUT/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 40 --warmup 3 --no-cpu --nbuf 4 > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ik_chain -s 8 -c 2 -o $OUT/prof_chain \
    python bench.py --steps 12 --warmup 3 --no-cpu --nbuf 4 > $OUT/ncu_full.log 2>&1
cp pink_b200/libpink_b200.so $OUT/libpink_b200.so
tail -5 $OUT/pytest_gpu.log; cat $OUT/smoke.log | tail -2; cat $OUT/bench.json; tail -3 $OUT/bench.err
