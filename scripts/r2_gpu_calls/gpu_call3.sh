#!/bin/bash
# round 2, GPU call 3 (2 GPUs): full suite incl. the multi-GPU tests, bench at N = 2 (C3 strong + C4 all-reduce leg)
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^\[BRIE2\]" > gpurun_out/r2_pytest_gpu_2gpu.log; grep -n "passed\|failed\|FAILED\|^C1\|^C2 batch\|n_iter\|Error" gpurun_out/r2_pytest_gpu_2gpu.log | cut -c1-600 | tail -60
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; tail -5 gpurun_out/r2_bench_2gpu.err; cat gpurun_out/r2_bench_2gpu.json
