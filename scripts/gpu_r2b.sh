#!/bin/bash
# round 2: A/B of the stencil kernel + a light ncu pass (stall reasons, icache, L1) on it
mkdir -p gpurun_out
TAG=${1:-r02b}
timeout 300 python scripts/ab_render.py > gpurun_out/${TAG}_ab.json 2> gpurun_out/${TAG}_ab.err
tail -3 gpurun_out/${TAG}_ab.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_ab.json"))
print({k:(round(v,3) if isinstance(v,float) else {a:round(b,6) for a,b in v.items()}) for k,v in d.items()})
PY
AC_RENDER_IMPL=st ncu --set full --clock-control none --import-source on -k regex:nsr_render_st -s 3 -c 1 -f -o gpurun_out/${TAG}_render \
    env AC_BENCH_SKIP_SDS=1 python bench.py --steps 1 --warmup 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
