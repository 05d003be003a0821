#!/bin/bash
# Config 5 sweep at the given rank counts (strong scaling, total batch 1K..1M). Usage: r2_sweep.sh tag "1 2"
TAG=${1:-r2s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for N in $2; do
  if [ $N -eq 1 ]; then
    timeout 600 python scripts/batch_sweep.py > $OUT/sweep_n1.jsonl 2> $OUT/sweep_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
        scripts/batch_sweep.py > $OUT/sweep_n$N.jsonl 2> $OUT/sweep_n$N.err
  fi
  echo "N=$N exit $?"; tail -2 $OUT/sweep_n$N.err | cut -c1-300
  python - <<PY
import json
for l in open("$OUT/sweep_n$N.jsonl"):
    if l.startswith("{"):
        d = json.loads(l)
        print("N=%d total %8d: solve %7.2f us  %.3e steps/s  hbm %.3f%s" % (d["n_gpus"], d["total_batch"], d["solve_us_per_step"], d["ik_steps_per_s"], d["hbm_frac_of_measured"],
              ("  +gather %7.2f us %.3e" % (d["solve_plus_gather_us_per_step"], d["solve_plus_gather_ik_steps_per_s"])) if "solve_plus_gather_us_per_step" in d else ""))
PY
done
