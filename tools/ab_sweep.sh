#!/bin/bash
# A/B of a build-time switch of the sweep kernel on the GPU box: tools/ab_sweep.sh "<flags A>" "<flags B>" ...
mkdir -p gpurun_out
for f in "$@"; do
  touch sfm-project_b200/csrc/match_tc.cu
  make -C sfm-project_b200/csrc EXTRA_match_tc="$f" > /dev/null 2>&1 || { echo "build failed for $f"; continue; }
  echo "=== flags: $f"
  timeout 200 python tools/bringup.py tc time 2>&1 | grep -E "identical|tcgen05:|sweep:"
  timeout 200 python tools/stage_times.py 2>/dev/null | tr -d '\n' | sed 's/"void at::native[^]]*\]//; s/"Mem[^]]*\],//g' | cut -c1-1400; echo
done
touch sfm-project_b200/csrc/match_tc.cu
make -C sfm-project_b200/csrc > /dev/null 2>&1
