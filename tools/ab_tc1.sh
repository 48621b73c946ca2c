#!/bin/bash
# A/B of build-time switches of the one-CTA sweep on the GPU box: tools/ab_tc1.sh <tag> "<flags A>" "<flags B>" ...
mkdir -p gpurun_out
TAG=$1; shift
for f in "$@"; do
  touch sfm-project_b200/csrc/match_tc.cu
  make -C sfm-project_b200/csrc EXTRA_match_tc="$f" > /dev/null 2>&1 || { echo "build failed for $f"; continue; }
  echo "=== flags: $f"
  timeout 200 python tools/bringup.py tc bounds time 2>&1 | grep -E "identical|differ|mismatch|mode [567]|tcgen05:|sweep:|probe|raised"
done 2>&1 | tee gpurun_out/${TAG}_ab_tc1.log
touch sfm-project_b200/csrc/match_tc.cu
make -C sfm-project_b200/csrc > /dev/null 2>&1
