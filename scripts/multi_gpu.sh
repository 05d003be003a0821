#!/bin/bash
# N-GPU bench exactly as the driver launches it.
N=${1:-2}
OUT=gpurun_out/${2:-mgpu}
mkdir -p $OUT
export PYTHONFAULTHANDLER=1
for n in ${ONLY_N-1} $N; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 5000 --warmup 20 --no-cpu > $OUT/bench_n1.json 2> $OUT/bench_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $n --steps 5000 --warmup 20 --no-cpu > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
    echo "torchrun exit $?" >> $OUT/bench_n$n.err
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29534 \
      bench.py --impl reference --gpus $n --steps 5 --warmup 1 > $OUT/ref_n$n.json 2> $OUT/ref_n$n.err
  fi
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_n$n.json")); print("N=$n value %.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], d.get("solve_plus_allgather"), d["gpu_launches"])
except Exception as e: print("N=$n ERR", e); print(open("$OUT/bench_n$n.err").read()[-2000:])
PY
done
tail -c 400 $OUT/ref_n$N.json
