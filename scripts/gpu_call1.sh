#!/bin/bash
# One gpurun call: GPU test suite, bench (both arms), train-step kernel breakdown, ncu launch list + full capture.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r01f_pytest.log 2>&1
tail -5 gpurun_out/r01f_pytest.log
timeout 600 python bench.py > gpurun_out/r01f_bench.json 2> gpurun_out/r01f_bench.err
cat gpurun_out/r01f_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01f_bench_reference.json 2>/dev/null
cat gpurun_out/r01f_bench_reference.json
timeout 300 python scripts/profile_train_step.py > gpurun_out/r01f_train_profile.txt 2>&1
tail -30 gpurun_out/r01f_train_profile.txt
timeout 600 bash scripts/profile_render.sh r01f
