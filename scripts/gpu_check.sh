#!/bin/bash
# Run on the GPU box (under gpurun): tests, bench, ncu launch list + full capture of the step kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests -m gpu -q 2>&1 | tail -15
python bench.py ${BENCH_ARGS:---steps 300 --warmup 10} > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ -n "$DO_NCU" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:elbo_step -s 4 -c 2 -f -o gpurun_out/prof \
      python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out
fi
