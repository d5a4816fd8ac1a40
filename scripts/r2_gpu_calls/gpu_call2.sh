#!/bin/bash
# round 2, GPU call 2 (1 GPU): full parity suite incl. the C1 fit, sanitizers, new bench.py at N = 1
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^\[BRIE2\]" | tail -80 > gpurun_out/r2_pytest_gpu.log; tail -30 gpurun_out/r2_pytest_gpu.log
TOOLS="memcheck racecheck" bash scripts/sanitize.sh
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; tail -5 gpurun_out/r2_bench_1gpu.err; cat gpurun_out/r2_bench_1gpu.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_ref.json 2>> gpurun_out/r2_bench_1gpu.err; cat gpurun_out/r2_bench_ref.json
