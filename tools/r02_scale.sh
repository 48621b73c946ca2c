#!/bin/bash
# sharded bench at N GPUs (strong scaling on configs[2]); usage: tools/r02_scale.sh <tag> <ngpus> [steps]
mkdir -p gpurun_out
TAG=${1:-r02q}; N=${2:-8}; STEPS=${3:-10}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
    --steps $STEPS --warmup 3 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
echo "bench n$N rc=$?"; python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/${TAG}_bench_n${N}.json") if l.startswith("{")][0]
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "scaling")}), json.dumps(d["config"]["sharded"]), json.dumps(d["config"]["selfcheck"]), json.dumps(d["config"]["one_gpu_same_workload"]), "e2e", d["e2e"]["value"], json.dumps(d["config"]["bank_broadcast_ms"]), json.dumps(d["clocks"]))
except Exception as e:
    print("no line:", e)
PY
tail -5 gpurun_out/${TAG}_bench_n${N}.err
