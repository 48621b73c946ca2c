#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02j}
timeout 240 python tools/r02_pair.py > gpurun_out/${TAG}_pair.log 2>&1; echo "pair rc=$?"; tail -8 gpurun_out/${TAG}_pair.log
bash tools/ab_pair_trace.sh ${TAG} > /dev/null 2>&1
grep -v "7[0-9]\{12\}" gpurun_out/${TAG}_pair_trace.log | cut -c1-220 | head -32
