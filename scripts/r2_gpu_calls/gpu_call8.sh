#!/bin/bash
# round 2, GPU call 8 (1 GPU, short): row-stride sweep (why did C2 slow down at ld 5120?), API / parity tests on HEAD
set -x
mkdir -p gpurun_out
for al in 32 64 128 256; do echo "align $al"; BRIE_LD_ALIGN=$al python scripts/scale_shapes.py C2 C3a --noloss 2>>gpurun_out/shapes.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['shape'], d['events'], 'kernel_ms', d['kernel_ms'], 'frac', d['frac_of_measured_hbm'])"; done
python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py tests/test_gpu_ingest.py tests/test_gpu_properties.py -m gpu -q -s 2>&1 | grep -v "^\[BRIE2\]" > gpurun_out/r2_pytest_gpu_call8.log; grep -n "passed\|failed\|FAILED\|violators\|Error" gpurun_out/r2_pytest_gpu_call8.log | cut -c1-400 | tail -30
