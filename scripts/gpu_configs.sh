#!/bin/bash
# On the GPU box: host probe, step-kernel shape scan, then full-size end-to-end config fits.
mkdir -p gpurun_out
{ nproc; free -g | head -2; df -h /tmp /dev/shm | tail -2; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; } > gpurun_out/host_probe.txt 2>&1
cat gpurun_out/host_probe.txt
if [ -n "$SHAPES" ]; then
  : > gpurun_out/shapes.jsonl
  for s in $SHAPES; do
    for v in noloss loss; do timeout 300 python scripts/scale_shapes.py $s --$v 2>&1 | tail -1 >> gpurun_out/shapes.jsonl; done
  done
  cat gpurun_out/shapes.jsonl | cut -c1-420
fi
for spec in "$@"; do
  echo "== config_fit $spec"
  BRIE_TIMING=1 timeout ${FIT_TIMEOUT:-900} python scripts/config_fit.py $spec > gpurun_out/fit.log 2> gpurun_out/fit.err
  echo "rc=$?"; tail -3 gpurun_out/fit.err; tail -2 gpurun_out/fit.log | cut -c1-1500
  tail -1 gpurun_out/fit.log >> gpurun_out/config_fits.jsonl
done
