#!/bin/bash
# Round-2 multi-GPU visit: the 2-GPU equality test, then the sharded bench (strong scaling on configs[2]).
# usage: tools/r02_multi.sh <tag> <ngpus> [steps] [transports]
mkdir -p gpurun_out
TAG=${1:-r02a}; N=${2:-2}; STEPS=${3:-5}; TRS=${4:-p2p}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.log 2>&1
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu -x -s > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "multi test rc=$?"; tail -40 gpurun_out/${TAG}_pytest_multi.log
for TR in $TRS; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
    --steps $STEPS --warmup 3 --transport $TR > gpurun_out/${TAG}_bench_n${N}_${TR}.json 2> gpurun_out/${TAG}_bench_n${N}_${TR}.err
echo "bench n$N $TR rc=$?"; python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/${TAG}_bench_n${N}_${TR}.json") if l.startswith("{")][0]
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}), json.dumps(d["config"]["sharded"]), json.dumps(d["config"]["selfcheck"]), json.dumps(d["config"]["one_gpu_same_workload"]), json.dumps(d["e2e"]["value"]))
except Exception as e:
    print("no line:", e)
PY
tail -5 gpurun_out/${TAG}_bench_n${N}_${TR}.err
done
