#!/bin/bash
mkdir -p gpurun_out
TAG=$1; shift
touch sfm-project_b200/csrc/match_tc2.cu
make -C sfm-project_b200/csrc EXTRA_match_tc2="-DSFM_TC2_TRACE=1 $*" > /dev/null 2>&1 || echo "build failed"
for m in 4 6; do echo "=== trace, sweep_only=$m"; timeout 120 python tools/r02_pair_trace.py $m 2>&1 | tail -32; done 2>&1 | tee gpurun_out/${TAG}_pair_trace.log
touch sfm-project_b200/csrc/match_tc2.cu
make -C sfm-project_b200/csrc > /dev/null 2>&1
