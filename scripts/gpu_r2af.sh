#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_warp.py tests/test_gpu_parity.py -q -x > gpurun_out/r02af_pytest.log 2>&1; tail -2 gpurun_out/r02af_pytest.log
AC_BENCH_SKIP_SDS=1 timeout 600 python bench.py > gpurun_out/r02af_bench.json 2> gpurun_out/r02af_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02af_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["warp_frame"].get("ms_per_frame"))
PY
