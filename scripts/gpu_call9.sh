#!/bin/bash
# round 2, GPU call 9 (1 GPU, short): HEAD check of bench.py (leg order changed) without the long e2e leg, smoke(), the resume test
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()"
python -m pytest tests/test_gpu_api.py -m gpu -q -k "resume or chunk or memmap or cli" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e > gpurun_out/r2_bench_1gpu_head_noe2e.json 2> gpurun_out/r2_bench_1gpu_head_noe2e.err; tail -3 gpurun_out/r2_bench_1gpu_head_noe2e.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_1gpu_head_noe2e.json').read().strip().splitlines()[-1])
print("value %.3e ms %.2f frac %.3f" % (d['value'], d['ms_per_step'], d['roofline']['frac']), d['clocks'])
print({k: (round(v['ms_per_step'], 3), round(v['frac_of_measured_hbm'], 3)) for k, v in d['shapes'].items()}, round(d['c4']['frac_of_measured_hbm'], 3))
PY
