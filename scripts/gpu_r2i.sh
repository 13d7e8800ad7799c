#!/bin/bash
# round 2: native VAE encoder forward + backward
mkdir -p gpurun_out
TAG=${1:-r02i}
timeout 900 python -m pytest tests/test_gpu_sds.py -q -s -x > gpurun_out/${TAG}_pytest_sds.log 2>&1
grep -n "passed\|failed\|rel L2\|Error\|error" gpurun_out/${TAG}_pytest_sds.log | head -40
tail -30 gpurun_out/${TAG}_pytest_sds.log
