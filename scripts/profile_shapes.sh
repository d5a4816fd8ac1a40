#!/bin/bash
# Run on the GPU box: one `ncu --set full` capture of the fused step kernel per shape/variant.
# usage: bash scripts/profile_shapes.sh "C3 C4" ; outputs gpurun_out/<shape>_<variant>.ncu-rep
mkdir -p gpurun_out
for s in ${1:-C3 C4}; do
  for v in noloss loss; do
    ncu --set full --clock-control none --import-source on -k regex:elbo_step -s 5 -c 1 -f \
        -o gpurun_out/${s}_${v} python scripts/scale_shapes.py $s --$v > gpurun_out/${s}_${v}.log 2>&1
    tail -2 gpurun_out/${s}_${v}.log
  done
done
ls -la gpurun_out
