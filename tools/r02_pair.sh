#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02d}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.log 2>&1
timeout 240 python tools/r02_pair.py > gpurun_out/${TAG}_pair.log 2>&1; echo "pair rc=$?"; tail -25 gpurun_out/${TAG}_pair.log
