bash tools/gpu_round.sh r01ac test smoke bench benchref launches full
timeout 600 python tools/time_verify.py > gpurun_out/r01ac_time_verify.json 2> gpurun_out/r01ac_time_verify.err; echo "time_verify rc=$?"
timeout 900 python tools/run_configs.py c3 c3m c4 c5 > gpurun_out/r01ac_configs_full.log 2>&1; echo "configs rc=$?"; tail -4 gpurun_out/r01ac_configs_full.log | cut -c1-400
timeout 300 python tools/stage_times.py > gpurun_out/r01ac_stages.json 2>/dev/null; echo "stages rc=$?"
