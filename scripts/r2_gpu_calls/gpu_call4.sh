#!/bin/bash
# round 2, GPU call 4 (1 GPU): target-run script at toy size, ncu captures of the HEAD kernels, sustained probe, wide shapes
set -x
mkdir -p gpurun_out
timeout 600 python scripts/target_run.py C3 --cells 20000 --events 600 --min-iter 600 --max-iter 2400 --n-eval 50 --json gpurun_out/r2_target_toy.json 2>&1 | grep -v "^\[BRIE2\]" | tail -5
timeout 600 python scripts/target_run.py C4 --cells 20000 --events 600 --min-iter 600 --max-iter 2400 --n-eval 50 2>&1 | grep -v "^\[BRIE2\]" | tail -3
python scripts/scale_shapes.py W16 W24 G20 K8G8 2>gpurun_out/shapes.err | tee gpurun_out/r2_shapes_wide.jsonl | cut -c1-400
bash scripts/profile_shapes.sh "C3 C4 W16" 2>&1 | tail -8
python scripts/sustained_probe.py C3 5 300 | tee gpurun_out/r2_sustained_C3.jsonl
python scripts/sustained_probe.py C2 5 3000 | tee gpurun_out/r2_sustained_C2.jsonl
python -m pytest tests/test_gpu_multi.py tests/test_gpu_api.py -m gpu -q 2>&1 | tail -3
