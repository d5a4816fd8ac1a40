#!/bin/bash
# rows-per-CTA scan on one box: bash scripts/rows_scan.sh C2 "128 136 144 200 256"
mkdir -p gpurun_out
python scripts/bw_probe.py
for rep in 1 2; do
  for r in $2; do
    BRIE_ROWS_PER_CTA=$r python scripts/scale_shapes.py $1 --noloss 2>> gpurun_out/rows.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('rows', d['rows_per_cta'], d['shape'], d['kernel_ms'], d['ms_per_step'], d['frac_of_measured_hbm'])"
  done
done
