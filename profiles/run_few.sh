#!/bin/bash
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "sars2_20k or many_clusters" > gpurun_out/${1:-few}_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/${1:-few}_pytest.log
