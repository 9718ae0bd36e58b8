#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mg.py tests/test_gpu_key_exchange.py -m gpu -x -q > gpurun_out/${1:-mgc}_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/${1:-mgc}_pytest.log
