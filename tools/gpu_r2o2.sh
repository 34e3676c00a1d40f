#!/bin/bash
# FP32 mean reversion in the difference form (hand-written Heston / OU, generated models): GPU suite, FP32 timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2o2.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_r2o2.log
tail -n 6 gpurun_out/pytest_r2o2.log
{
for r in 1 2; do echo -n "f32 bench: "; SDEB200_RNG=2 python tools/prof_sample.py f32 19398656 bench 2>&1 | tail -n 1 | awk '{print $(NF-3), $(NF-2), $(NF-1)}'; done
echo -n "f64 bench: "; SDEB200_RNG=2 python tools/prof_sample.py f64 19398656 bench 2>&1 | tail -n 1 | awk '{print $(NF-3), $(NF-2), $(NF-1)}'
SDEB200_RNG=2 timeout 600 python tools/prof_generated.py 2>&1 | tail -n 8
} | tee gpurun_out/r02zz_fp32_mean_reversion.txt
