#!/bin/bash
# A/B runs of bench.py under different environment settings: profiles/ab_env.sh <tag> "<ENV1=..>" "<ENV2=.. ENV3=..>" ...
tag=$1; shift
mkdir -p gpurun_out
i=0
for v in "$@"; do
  i=$((i+1))
  env $v python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  echo "[$v] rc=$?"; tail -2 gpurun_out/${tag}_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_$i.json")); print("[$v]", round(d["ms_per_step"],2), {k:round(x,2) for k,x in d["roofline"]["kernels_ms"].items()})
except Exception as e: print("no json", e)
PY
done
