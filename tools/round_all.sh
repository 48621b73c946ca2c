#!/bin/bash
# Everything the round's profiles/ are built from, on one box: tools/round_all.sh <tag>
TAG=${1:-r01x}
bash tools/gpu_round.sh $TAG test smoke bench benchref launches full
timeout 600 python tools/time_verify.py > gpurun_out/${TAG}_time_verify.json 2> gpurun_out/${TAG}_time_verify.err; echo "time_verify rc=$?"
timeout 300 python tools/stage_times.py > gpurun_out/${TAG}_stages.json 2>/dev/null; echo "stages rc=$?"
timeout 300 python tools/prof_verify.py 2 > gpurun_out/${TAG}_plain_verify.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ransac_f_kernel|ransac_h_kernel|pose_kernel" -s 4 -c 4 -o gpurun_out/${TAG}_prof_verify -f python tools/prof_verify.py 2 > gpurun_out/${TAG}_ncu_verify.log 2>&1; echo "prof_verify rc=$?"
timeout 1200 python tools/run_configs.py c3 c3m c4 c5 > gpurun_out/${TAG}_configs_full.log 2>&1; echo "configs rc=$?"
timeout 300 python tools/e2e_exp.py > gpurun_out/${TAG}_e2e_exp.log 2>&1; echo "e2e_exp rc=$?"
