#!/bin/bash
# Runs every bring-up stage in its own process with a timeout; log -> gpurun_out/bringup.log
mkdir -p gpurun_out
LOG=gpurun_out/bringup.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
for s in ${@:-pack simt tile1 tile2 tile0 tc filter hamming ransac time}; do
  timeout 240 python tools/bringup.py $s >> $LOG 2>&1
  echo "--- exit $? for $s" >> $LOG
done
tail -120 $LOG
