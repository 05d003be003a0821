#!/bin/bash
# e2e throughput vs submission streams and chunk size.  Usage: bash scripts/e2e_sweep.sh [tag]
OUT=gpurun_out/${1:-e2e}
mkdir -p $OUT
for DUP in 0 1; do
 for NS in 1 2; do
  for CH in 8192 16384 32768; do
    PK_HOST_DUPLEX=$DUP PK_HOST_CHUNK=$CH timeout 300 python bench.py --steps 2000 --warmup 20 --no-cpu --e2e-streams $NS 2>/dev/null \
      | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('duplex', $DUP, 'streams', $NS, 'chunk', $CH, 'e2e', d['e2e']['value'], 'us', d['e2e']['ms_per_step']*1e3, 'ok', d['e2e']['bitwise_equal_to_device_path'])" | tee -a $OUT/sweep.txt
  done
 done
done
python scripts/pcie_probe.py > $OUT/pcie.txt 2>&1; tail -3 $OUT/pcie.txt
