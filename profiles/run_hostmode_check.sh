#!/bin/bash
tag=${1:-hm}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py tests/test_gpu_mg.py tests/test_packed.py -m gpu -x -q -k "host or config5 or many_breaks or gff or empty or mg_ or packed or flag_combinations" > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
python bench.py --config 5 --steps 3 --warmup 2 > gpurun_out/${tag}_config5.json 2> gpurun_out/${tag}_config5.err; echo "config5 rc=$?"; tail -2 gpurun_out/${tag}_config5.err
