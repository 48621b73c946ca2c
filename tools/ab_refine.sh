#!/bin/bash
# same-box A/B of build-time switches on the fused refinement: tools/ab_refine.sh <tag> "<flags A>" "<flags B>" ...
mkdir -p gpurun_out
TAG=$1; shift
for f in "$@"; do
  touch sfm-project_b200/csrc/match_tc.cu
  make -C sfm-project_b200/csrc EXTRA_match_tc="$f" > /dev/null 2>&1 || { echo "build failed for $f"; continue; }
  echo "=== flags: $f"
  timeout 200 python tools/stage_times.py 2>/dev/null | python -c "
import sys, json
txt = sys.stdin.read()
dec = json.JSONDecoder(); i = 0; objs = []
while i < len(txt):
    while i < len(txt) and txt[i].isspace(): i += 1
    if i >= len(txt): break
    o, j = dec.raw_decode(txt, i); objs.append(o); i = j
print('whole step', objs[0]['whole_step_ms'], 'sweep', objs[0]['sweep_ms'])
for k, v in objs[1].items():
    if 'refine' in k or 'ransac' in k or 'match_tc' in k: print(' ', k[:50], v)
"
done 2>&1 | tee gpurun_out/${TAG}_ab_refine.log
touch sfm-project_b200/csrc/match_tc.cu
make -C sfm-project_b200/csrc > /dev/null 2>&1
