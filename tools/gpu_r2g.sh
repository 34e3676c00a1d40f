#!/bin/bash
# FAST forms of generated functors (folded constants, factored state, expanded mean reversion): GPU suite on the variant build
# $1 (default: affine), then generated vs built-in
V=${1:-affine}
mkdir -p gpurun_out
export SDEB200_LIB=$PWD/variants/libsdeb200_$V.so
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2g.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_r2g.log
tail -n 8 gpurun_out/pytest_r2g.log
for r in 0 2; do echo "== generated vs built-in, SDEB200_RNG=$r"; SDEB200_RNG=$r timeout 600 python tools/prof_generated.py 2>&1 | tail -n 8; done | tee gpurun_out/generated_r2g_$V.txt
