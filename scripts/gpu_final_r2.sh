#!/bin/bash
# round 2 final check: smoke(), full GPU suite, bench (the driver's sequence)
mkdir -p gpurun_out
TAG=${1:-r02F}
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_all.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_all.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/${TAG}_bench_reference.json
timeout 1200 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","steps","warmup")}, d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["clocks"])
s=d["sds_step"]; print(s.get("value"), s.get("ms_per_step"), s.get("phases_ms"), s.get("error"))
print(s["nerf_side_only"]["ms_per_step"], s["coarse_stage_nerf_side_only"]["ms_per_step"], s.get("reference_gpu_path",{}).get("train_step_nerf_side_ms"), s["config5_512_multibbox_sd21"].get("ms_per_step"))
print(d["warp_frame"].get("ms_per_frame"))
PY
