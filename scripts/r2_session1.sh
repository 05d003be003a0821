#!/bin/bash
# Round-2 GPU session 1: GPU tests (incl. the ones added after round 1's last GPU run), bench at the
# driver's settings and at a long run, e2e schedules, reference arm, microbenchmarks.
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
nvidia-smi topo -m >> $OUT/gpu.txt 2>&1
(nproc; lscpu | head -25; cat /sys/devices/system/node/online; for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist); done) > $OUT/host.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_20.json 2> $OUT/bench_20.err
echo "bench exit $?" >> $OUT/bench_20.err
timeout 600 python bench.py --steps 2000 --warmup 20 --no-cpu --no-configs > $OUT/bench_2000.json 2> $OUT/bench_2000.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --no-numa-bind > $OUT/bench_20_nonuma.json 2>> $OUT/bench_20.err
PK_HOST_DUPLEX=1 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --no-configs > $OUT/bench_duplex.json 2>> $OUT/bench_20.err
PK_HOST_MODE=2 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --no-configs > $OUT/bench_mode2.json 2>> $OUT/bench_20.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2>> $OUT/bench_20.err
timeout 120 scripts/ubench/ubench > $OUT/ubench.txt 2>&1
timeout 120 python scripts/pcie_probe.py > $OUT/pcie.txt 2>&1
tail -3 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; cat $OUT/bench_20.json | head -c 3000; echo; tail -3 $OUT/bench_20.err; cat $OUT/ubench.txt
