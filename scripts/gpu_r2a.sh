#!/bin/bash
# round 2, call A: new stencil kernel -- A/B against the round-1 kernel, GPU suite, bench
mkdir -p gpurun_out
timeout 300 python scripts/ab_render.py > gpurun_out/r02a_ab.json 2> gpurun_out/r02a_ab.err
tail -5 gpurun_out/r02a_ab.err
cat gpurun_out/r02a_ab.json
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02a_pytest.log 2>&1
tail -15 gpurun_out/r02a_pytest.log
