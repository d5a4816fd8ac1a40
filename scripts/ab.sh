#!/bin/bash
# A/B of library variants on ONE box (box-to-box variance is a few %):
#   bash scripts/ab.sh "ring2 ring3" "C2 C3 C4 C5" [repeats]
# variants are brie_b200/variants/<name>.so (built with BRIE_LIB_OUT / BRIE_NVCC_EXTRA)
mkdir -p gpurun_out
python scripts/bw_probe.py | tee -a gpurun_out/ab.jsonl
for rep in $(seq 1 ${3:-2}); do
  for v in $1; do
    BRIE_LIB_PATH=$PWD/brie_b200/variants/$v.so python scripts/scale_shapes.py $2 2>> gpurun_out/ab.err \
      | sed "s/^{/{\"variant\": \"$v\", \"rep\": $rep, /" | tee -a gpurun_out/ab.jsonl \
      | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['variant'], d['rep'], d['shape'], 'loss' if d['loss_trace'] else 'noloss', d['kernel_ms'], d['frac_of_measured_hbm'])"
  done
done
