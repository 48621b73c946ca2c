#!/bin/bash
# A/B of build-time switches of the CTA-pair sweep on the GPU box: tools/ab_pair.sh <tag> "<flags A>" "<flags B>" ...
mkdir -p gpurun_out
TAG=$1; shift
for f in "$@"; do
  touch sfm-project_b200/csrc/match_tc2.cu
  make -C sfm-project_b200/csrc EXTRA_match_tc2="$f" > /dev/null 2>&1 || { echo "build failed for $f"; continue; }
  echo "=== flags: $f"
  timeout 120 python tools/r02_pair.py 2>&1 | tail -7
done 2>&1 | tee gpurun_out/${TAG}_ab_pair.log
touch sfm-project_b200/csrc/match_tc2.cu
make -C sfm-project_b200/csrc > /dev/null 2>&1
