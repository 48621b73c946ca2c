#!/bin/bash
# N-GPU sharded bench with the host timeline of every rank (SFM_HOST_TRACE=1): who waits for whom, step by step
mkdir -p gpurun_out
N=${1:-2}; REPS=${2:-3}
for i in $(seq 1 $REPS); do
SFM_HOST_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 --transport p2p > gpurun_out/ht_tmp.json 2> gpurun_out/ht_tmp.err
python - <<PY
import json
ev = [json.load(open(f"gpurun_out/hosttrace_rank{r}.json")) for r in range($N)]
w0 = ev[0][0]["wall"][0]
d = [json.loads(l) for l in open("gpurun_out/ht_tmp.json") if l.startswith("{")][0]
print("run $i: value", round(d["value"]), "ms/step", round(d["ms_per_step"], 2))
for s in range(len(ev[0])):
    row = []
    for r in range($N):
        e = ev[r][s]
        a, b, c = [1e3 * (x - w0) for x in e["wall"]]
        row.append(f"r{r}: start {a:8.1f} ret {b:8.1f} sync {c:8.1f} | gpu compute {e['t_compute']:6.1f} fence {e.get('t_fence') or 0:6.1f} done {e['t_done']:6.1f}")
    print(f"  step {s}  " + "   ".join(row))
PY
done 2>&1 | tee gpurun_out/r02_hosttrace.log
