#!/bin/bash
# round 2, GPU call 12 (1 GPU): the whole `-m gpu` suite on the final HEAD, as the driver runs it
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2_pytest_gpu_head_final_tail.log
