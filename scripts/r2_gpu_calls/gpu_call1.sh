#!/bin/bash
# round 2, GPU call 1: parity suite with the new BASELINE-config cases, sanitizers, A/B of the packed-f32 build
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc; free -g | head -2; df -h / /dev/shm | tail -2
python -m pytest tests -m "gpu and not slow" -q -x 2>&1 | tail -25 > gpurun_out/r2_pytest_gpu.log; tail -8 gpurun_out/r2_pytest_gpu.log
python -m pytest tests/test_gpu_baseline_parity.py -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r2_pytest_baseline.log; tail -40 gpurun_out/r2_pytest_baseline.log
bash scripts/ab.sh "head f2" "C2 C3 C4 W16" 2 2>&1 | tail -40
BRIE_LIB_PATH=$PWD/brie_b200/variants/f2.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5
TOOLS="memcheck racecheck" bash scripts/sanitize.sh
