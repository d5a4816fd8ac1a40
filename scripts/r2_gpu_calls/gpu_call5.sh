#!/bin/bash
# round 2, GPU call 5 (8 GPUs): the north_star target run (C5 through fitBRIE on 8 x B200), full C4 with the in-library
# all-reduce, bench.py at N = 8 and N = 4
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2; df -h /dev/shm /tmp | tail -2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 1500 $TR --nproc-per-node 8 --master-port 29721 scripts/target_run.py C5 --json gpurun_out/r2_target_C5_8gpu.json 2>&1 | grep -v "^\[BRIE2\]" | tail -4 | cut -c1-3000
rm -rf /dev/shm/brie_target_C5
timeout 600 $TR --nproc-per-node 8 --master-port 29722 scripts/target_run.py C4 --json gpurun_out/r2_target_C4_8gpu.json 2>&1 | grep -v "^\[BRIE2\]" | tail -4 | cut -c1-3000
rm -rf /dev/shm/brie_target_C4
timeout 800 $TR --nproc-per-node 8 --master-port 29723 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; tail -3 gpurun_out/r2_bench_8gpu.err; cut -c1-1500 gpurun_out/r2_bench_8gpu.json
rm -rf /dev/shm/brie_bench_e2e_C3
timeout 800 $TR --nproc-per-node 4 --master-port 29724 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err; tail -3 gpurun_out/r2_bench_4gpu.err; cut -c1-600 gpurun_out/r2_bench_4gpu.json
