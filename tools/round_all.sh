bash tools/gpu_round.sh r01ad test smoke bench benchref launches full
timeout 600 python tools/time_verify.py > gpurun_out/r01ad_time_verify.json 2> gpurun_out/r01ad_time_verify.err; echo "time_verify rc=$?"
timeout 300 python tools/stage_times.py > gpurun_out/r01ad_stages.json 2>/dev/null; echo "stages rc=$?"
timeout 300 python tools/prof_verify.py 2 > gpurun_out/r01ad_plain_verify.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ransac_f_kernel|ransac_h_kernel|pose_kernel" -s 4 -c 4 -o gpurun_out/r01ad_prof_verify -f python tools/prof_verify.py 2 > gpurun_out/r01ad_ncu_verify.log 2>&1; echo "prof_verify rc=$?"
