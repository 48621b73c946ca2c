#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02ad}
for V in "--transport p2p" "--transport p2p --no-overlap" "--transport sendrecv" "--transport p2p --gather summaries"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 2 $V > gpurun_out/${TAG}_tmp.json 2> gpurun_out/${TAG}_tmp.err
echo "== $V rc=$?"; python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/${TAG}_tmp.json") if l.startswith("{")][0]
    s = d["config"]["sharded"]
    print(round(d["value"]), round(d["ms_per_step"], 2), "slow", round(s["compute_ms_slowest_rank"], 2), "fast", round(s["compute_ms_fastest_rank"], 2), "rank0", round(s["compute_ms_rank0"], 2), "exposed", round(s["gather_ms_exposed_on_rank0"], 2))
except Exception as e:
    print("no line:", e)
PY
done 2>&1 | tee gpurun_out/${TAG}_n2_diag.log
