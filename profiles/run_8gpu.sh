#!/bin/bash
# 8-GPU box: library collectives check (in-process npt_mg_*), headline bench, ingest ceiling, configs 3 and 5
tag=${1:-mg8}; N=${2:-8}
mkdir -p gpurun_out
devs=$(seq -s, 0 $((N-1)))
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
python tests/mg_check.py --devices $devs 2>&1 | grep -E "mg_check|Error|error" | tail -3
$TR --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${tag}_bench.err
$TR --master-port 29522 profiles/h2d_ceiling.py 2>&1 | grep h2d_ceiling | tee gpurun_out/${tag}_h2d_ceiling.txt
$TR --master-port 29523 bench.py --gpus $N --config 3 --steps 3 --warmup 2 > gpurun_out/${tag}_config3.json 2> gpurun_out/${tag}_config3.err; echo "config3 rc=$?"; tail -2 gpurun_out/${tag}_config3.err
$TR --master-port 29524 bench.py --gpus $N --config 5 --steps 3 --warmup 2 > gpurun_out/${tag}_config5.json 2> gpurun_out/${tag}_config5.err; echo "config5 rc=$?"; tail -2 gpurun_out/${tag}_config5.err
