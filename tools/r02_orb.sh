#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02p}
timeout 600 python -m pytest tests/test_gpu_orb.py tests/test_gpu_matcher.py -q -m gpu -x -s > gpurun_out/${TAG}_pytest_orb.log 2>&1; echo "pytest rc=$?"; grep -E "ms per image|passed|failed|Error|assert" gpurun_out/${TAG}_pytest_orb.log | head -20
