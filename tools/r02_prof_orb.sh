#!/bin/bash
# ORB extraction: parity tests, then the host timeline / kernel launch list of one 1080p image
mkdir -p gpurun_out
TAG=${1:-r02ap}
timeout 600 python -m pytest tests/test_gpu_orb.py -q -m gpu -x -s > gpurun_out/${TAG}_pytest_orb.log 2>&1; echo "pytest rc=$?"; grep -E "ms per image|passed|failed|Error|assert" gpurun_out/${TAG}_pytest_orb.log | head -20
timeout 300 python tools/prof_orb.py > gpurun_out/${TAG}_prof_orb.log 2>&1; echo "rc=$?"
head -12 gpurun_out/${TAG}_prof_orb.log
