#!/bin/bash
# Multi-GPU session: peer-gather check and the bench line (driver launch) at N ranks.
# Usage: bash scripts/r2_multi.sh tag N [steps]
TAG=${1:-r2e}
N=${2:-2}
STEPS=${3:-20}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_peer_gather.py -m gpu -q > $OUT/pytest_peer.log 2>&1; tail -2 $OUT/pytest_peer.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    scripts/peer_gather_check.py > $OUT/peer_check.log 2>&1
echo "peer check exit $?"; tail -3 $OUT/peer_check.log
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=$OUT/nccl_%h_%p.log timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps $STEPS --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "bench exit $?"; tail -3 $OUT/bench_n$N.err
python - <<PY
import json
d = json.load(open("$OUT/bench_n$N.json"))
print("N=%d value %.3e (%.2f us/step) e2e %.3e (%.1f us/step)" % (d["n_gpus"], d["value"], d["ms_per_step"]*1e3, d["e2e"]["value"], d["e2e"]["ms_per_step"]*1e3))
print(json.dumps(d.get("solve_plus_allgather"), indent=1))
PY
grep -h "nranks" $OUT/nccl_*.log | head -3
