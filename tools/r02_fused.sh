#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02t}
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python tools/stage_times.py > gpurun_out/${TAG}_stages.json 2> gpurun_out/${TAG}_stages.err; echo "stages rc=$?"; cut -c1-1500 gpurun_out/${TAG}_stages.json; tail -3 gpurun_out/${TAG}_stages.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_n1.json; tail -3 gpurun_out/${TAG}_bench_n1.err
