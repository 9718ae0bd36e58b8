#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${1:-suite}_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/${1:-suite}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
