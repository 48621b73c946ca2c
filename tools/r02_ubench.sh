#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02c}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.log 2>&1
timeout 120 tools/ubench/sweep_parts > gpurun_out/${TAG}_sweep_parts.log 2>&1; echo "sweep_parts rc=$?"; cat gpurun_out/${TAG}_sweep_parts.log
