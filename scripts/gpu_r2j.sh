#!/bin/bash
# round 2: bench after the native VAE (sds_step phases) + torch-profiler kernel table of the guidance
mkdir -p gpurun_out
TAG=${1:-r02j}
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"])
s=d["sds_step"]; print(s.get("value"), s.get("ms_per_step"), s.get("phases_ms"), s.get("error"))
print(s.get("nerf_side_only")); print(s.get("coarse_stage_nerf_side_only"))
PY
timeout 600 python scripts/profile_guidance.py > gpurun_out/${TAG}_sds_profile.txt 2>&1
tail -45 gpurun_out/${TAG}_sds_profile.txt
