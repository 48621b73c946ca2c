#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
for i in 1 2 3; do
SFM_HOST_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 --transport p2p > gpurun_out/ht_tmp.json 2> gpurun_out/ht_tmp.err
python - <<PY
import json
for r in range($N):
    ev = json.load(open(f"gpurun_out/hosttrace_rank{r}.json"))
    print("run $i rank", r, "compute", [round(e["t_compute"], 1) for e in ev], "host", [e["host_ms"] for e in ev])
    worst = max(ev, key=lambda e: e["t_compute"])
    tr = worst["host_trace_ms"]
    tri = [(round(tr[i + 1] - tr[i], 2), round(tr[i + 2] - tr[i + 1], 2)) for i in range(0, len(tr) - 2, 3)]
    print("   worst step", round(worst["t_compute"], 1), "per batch (launch ms, push ms):", tri, "last stamp", tr[-1])
PY
done 2>&1 | tee gpurun_out/r02ak_hosttrace.log
