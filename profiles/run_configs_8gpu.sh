#!/bin/bash
tag=${1:-cfg8}; N=${2:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29523 bench.py --gpus $N --config 3 --steps 3 --warmup 2 > gpurun_out/${tag}_config3.json 2> gpurun_out/${tag}_config3.err; echo "config3 rc=$?"; grep -E "NptError|Error:" gpurun_out/${tag}_config3.err | head -3
$TR --master-port 29525 bench.py --gpus $N --config 4 --steps 2 --warmup 1 > gpurun_out/${tag}_config4.json 2> gpurun_out/${tag}_config4.err; echo "config4 rc=$?"; grep -E "NptError|Error:" gpurun_out/${tag}_config4.err | head -3
