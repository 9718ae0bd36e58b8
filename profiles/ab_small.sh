#!/bin/bash
# quick A/B of library builds: profiles/ab_small.sh <tag> "<ENV ...>" ...   (4 M reads, no CPU leg)
tag=$1; shift
mkdir -p gpurun_out
i=0
for v in "$@"; do
  i=$((i+1))
  env $v python bench.py --steps 6 --warmup 3 --no-cpu --e2e-steps 1 --e2e-chunks 1 --reads 4000000 > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  echo "[$v] rc=$?"; tail -1 gpurun_out/${tag}_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_$i.json")); print("[$v]", round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["kernels_ms"].items()}, d["config"]["self_check"]["passed"].__len__())
except Exception as e: print("no json", e)
PY
done
