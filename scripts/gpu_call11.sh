#!/bin/bash
# round 2, GPU call 11 (1 GPU, short): sustained probes on HEAD (512-byte aligned rows)
set -x
mkdir -p gpurun_out
python scripts/sustained_probe.py C2 3 3000 | tee gpurun_out/r2_sustained_head.jsonl
python scripts/sustained_probe.py C3 3 300 | tee -a gpurun_out/r2_sustained_head.jsonl
python scripts/sustained_probe.py W16 3 1500 | tee -a gpurun_out/r2_sustained_head.jsonl
python scripts/sustained_probe.py C4 3 300 | tee -a gpurun_out/r2_sustained_head.jsonl
