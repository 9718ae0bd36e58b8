#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bam_file" > gpurun_out/${1:-one}_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/${1:-one}_pytest.log
