#!/bin/bash
# round 2: tightened parity gates, SH vs reference kernel, warp tie check -- full GPU suite with the measured fractions printed
mkdir -p gpurun_out
TAG=${1:-r02f}
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest_all.log 2>&1
grep -n "passed\|failed\|identical\|agree with the reference\|coincidence\|SH degree\|closest-point\|relative L2\|Error" gpurun_out/${TAG}_pytest_all.log | head -80
