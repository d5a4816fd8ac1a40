#!/bin/bash
# round 2, GPU call 7 (1 GPU): HEAD verification -- the whole GPU suite (incl. the C1 fit), sanitizers over the new kernel
# forms, bench.py at N = 1 with the 512-byte aligned leading dimension
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^\[BRIE2\]" > gpurun_out/r2_pytest_gpu_final.log; grep -n "passed\|failed\|FAILED\|^C1\|^C2 batch\|Error" gpurun_out/r2_pytest_gpu_final.log | cut -c1-500 | tail -40
TOOLS="memcheck racecheck" bash scripts/sanitize.sh
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu_final.json 2> gpurun_out/r2_bench_1gpu_final.err; tail -5 gpurun_out/r2_bench_1gpu_final.err; cut -c1-2500 gpurun_out/r2_bench_1gpu_final.json
