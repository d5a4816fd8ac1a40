#!/bin/bash
# round 2, GPU call 10 (2 GPUs): multi-GPU tests and bench.py at N = 2 on HEAD (512-byte aligned rows, background D2H)
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -q -s 2>&1 | grep -v "^\[BRIE2\]" > gpurun_out/r2_pytest_multi_head.log; tail -4 gpurun_out/r2_pytest_multi_head.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_2gpu_head.json 2> gpurun_out/r2_bench_2gpu_head.err; tail -3 gpurun_out/r2_bench_2gpu_head.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_2gpu_head.json').read().strip().splitlines()[-1])
print("value %.3e ms %.2f frac %.3f e2e %.3e wall %.1f" % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['wall_s']), d['clocks'])
print({k: (round(v['ms_per_step'], 3), round(v['frac_of_measured_hbm'], 3)) for k, v in d['shapes'].items()})
print({k: v for k, v in d['c4'].items() if k not in ('workload', 'collective')})
PY
