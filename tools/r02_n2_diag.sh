#!/bin/bash
# N-GPU sharded bench, repeated, with the per-rank split (kernels / pushes / clocks): tools/r02_n2_diag.sh <tag> <ngpus> "<variant flags>" ...
mkdir -p gpurun_out
TAG=${1:-r02ad}; N=${2:-2}; shift; shift
nvidia-smi topo -m > gpurun_out/${TAG}_topo.log 2>&1
for V in "$@"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 $V > gpurun_out/${TAG}_tmp.json 2> gpurun_out/${TAG}_tmp.err
echo "== $V rc=$?"; python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/${TAG}_tmp.json") if l.startswith("{")][0]
    s = d["config"]["sharded"]
    print(round(d["value"]), round(d["ms_per_step"], 2), "exposed", round(s["gather_ms_exposed_on_rank0"], 2), "e2e", round(d["e2e"]["value"]))
    for r, x in enumerate(s["per_rank"]):
        print("  rank", r, {k: v for k, v in x.items()})
    print("  clocks", d["clocks"].get("per_rank_sm_mhz"), d["clocks"].get("per_rank_power_w_max"), d["clocks"]["reasons"])
except Exception as e:
    print("no line:", e)
PY
done 2>&1 | tee gpurun_out/${TAG}_n${N}_diag.log
