#!/bin/bash
# one GPU: the -m gpu suite, the headline bench line, every other configuration, then the ncu launch list and one
# --set full capture of the per-batch kernels at 4 M reads per launch (numbers printed under ncu are never bench values)
tag=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference rc=$?"
for c in 1 3 4 5; do
  python bench.py --config $c --steps 3 --warmup 2 > gpurun_out/${tag}_config$c.json 2> gpurun_out/${tag}_config$c.err; echo "config $c rc=$?"; tail -2 gpurun_out/${tag}_config$c.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/${tag}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"fast_walk_kernel|accumulate_kernel|cluster_kernel|cov_key_kernel" -s 4 -c 4 -o gpurun_out/${tag} python bench.py --steps 1 --warmup 1 --reads 4000000 --no-cpu --e2e-steps 1 > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
