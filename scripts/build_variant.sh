#!/bin/bash
# Build a library variant for scripts/ab.sh:  bash scripts/build_variant.sh <name> ["extra nvcc flags"]
#   bash scripts/build_variant.sh head                      # current source, default flags
#   bash scripts/build_variant.sh f2 "-DBRIE_F32X2=1"       # packed-FP32 form of the dense phases
# then on the GPU box:  bash scripts/ab.sh "head f2" "C2 C3 C4 C5 W16" 1
# (parity of a variant: BRIE_LIB_PATH=$PWD/brie_b200/variants/f2.so python -m pytest tests -m gpu -q)
set -e
mkdir -p brie_b200/variants
BRIE_LIB_OUT=$PWD/brie_b200/variants/$1.so BRIE_NVCC_EXTRA="$2" python -m brie_b200.build --force
ls -la brie_b200/variants/$1.so
