#!/bin/bash
# correctness of the sweep as built (matcher tests), then the build-time A/B of the epilogue form on the same box
TAG=${1:-r02ae}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_matcher.py -x -q -m gpu > gpurun_out/${TAG}_pytest_matcher.log 2>&1
echo "matcher tests rc=$?"; tail -5 gpurun_out/${TAG}_pytest_matcher.log
bash tools/ab_tc1.sh ${TAG} "-DSFM_TC_EPI=2" "-DSFM_TC_EPI=1" "-DSFM_TC_EPI=2 -DSFM_TC_SEQ=0"
