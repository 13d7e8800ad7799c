#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02n}
timeout 600 python -m pytest tests/test_gpu_frame_ops.py tests/test_gpu_warp.py -q -s -k "iso_surface or warped_render or entry_scripts" > gpurun_out/${TAG}_pytest.log 2>&1
grep -n "passed\|failed\|iso-surface\|Error" gpurun_out/${TAG}_pytest.log | head
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -2 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"])
s=d["sds_step"]; print(s.get("value"), s.get("ms_per_step"), s.get("phases_ms"), s.get("error"))
print(s.get("config5_512_multibbox_sd21"))
PY
