#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/${1:-b}_bench.json 2> gpurun_out/${1:-b}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${1:-b}_bench.err
