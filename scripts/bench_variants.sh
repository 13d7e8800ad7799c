#!/bin/bash
# Time every tuning variant of the library built under avatarcraft_b200/_variants/ (one gpurun call).
for f in avatarcraft_b200/_variants/*.so; do
  echo "== $f"
  AC_LIB_PATH=$PWD/$f timeout 200 python bench.py --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
done
