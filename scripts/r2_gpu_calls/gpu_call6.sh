#!/bin/bash
# round 2, GPU call 6 (1 GPU): A/B of the two-rows-per-iteration form and the TMA-bulk ring; full parity suite on HEAD;
# ncu digest of the full-size C3 launch (why 0.89 at 10 000 events when the 2 048-event slab runs at 0.94)
set -x
mkdir -p gpurun_out
python -m pytest tests -m "gpu and not slow" -q 2>&1 | tail -4
BRIE_LIB_PATH=$PWD/brie_b200/variants/tma.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "first_step or trajectory or loss_gene" 2>&1 | tail -3
timeout 600 bash scripts/ab.sh "head rpi1 tma" "C2 C3 W16 K8G8" 1 2>&1 | grep -v "^+" | tail -70
python scripts/scale_shapes.py C3 C3a C3b C3F W24 G20 --noloss 2>>gpurun_out/shapes.err | cut -c1-500
ncu --set full --clock-control none --import-source on -k regex:elbo_step -s 5 -c 1 -f -o gpurun_out/C3F_noloss python scripts/scale_shapes.py C3F --noloss > gpurun_out/C3F_noloss.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:elbo_step -s 5 -c 1 -f -o gpurun_out/W16_noloss python scripts/scale_shapes.py W16 --noloss > gpurun_out/W16_noloss.log 2>&1
python scripts/ncu_digest.py gpurun_out/C3F_noloss.ncu-rep gpurun_out/W16_noloss.ncu-rep > gpurun_out/r2_ncu_digest.md 2>&1
for s in C3F W16; do python scripts/ncu_hot.py gpurun_out/${s}_noloss.ncu-rep 30 > gpurun_out/r2_ncu_hot_${s}.txt 2>&1; done
rm -f gpurun_out/W16_noloss.ncu-rep
ls -la gpurun_out; head -60 gpurun_out/r2_ncu_digest.md
