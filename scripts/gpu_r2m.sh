#!/bin/bash
# round 2: full GPU suite + bench after the split backward / batched sampling
mkdir -p gpurun_out
TAG=${1:-r02m}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_all.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_all.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"])
s=d["sds_step"]; print(s.get("value"), s.get("ms_per_step"), s.get("phases_ms"), s.get("error"))
print(s.get("nerf_side_only")); print(s.get("coarse_stage_nerf_side_only")); print(s.get("reference_gpu_path"))
print(d["warp_frame"])
PY
