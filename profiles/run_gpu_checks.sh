#!/bin/bash
# GPU box: the -m gpu suite, the default bench line and the host-mode config, outputs under gpurun_out/<tag>_*
tag=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -4 gpurun_out/${tag}_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${tag}_bench.err
python bench.py --config 5 --steps 3 --warmup 2 > gpurun_out/${tag}_config5.json 2> gpurun_out/${tag}_config5.err; echo "config5 rc=$?"; tail -2 gpurun_out/${tag}_config5.err
