#!/bin/bash
# round 2: failed tests of r02f re-run, then the full bench (render + sds_step with phases + reference GPU path + warp frame)
mkdir -p gpurun_out
TAG=${1:-r02g}
timeout 900 python -m pytest tests/test_gpu_encoders.py tests/test_gpu_warp.py tests/test_gpu_parity.py -q -s > gpurun_out/${TAG}_pytest.log 2>&1
grep -n "passed\|failed\|SH degree\|unit directions\|closest\|differing\|agree with\|Error" gpurun_out/${TAG}_pytest.log | head -50
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"])
print(json.dumps(d["sds_step"], indent=1)[:3500])
print(d["warp_frame"])
PY
