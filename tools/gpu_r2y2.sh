#!/bin/bash
mkdir -p gpurun_out
VARIANTS="cur v2ppt4 p4u2 p2u2 p4mb3 p2mb3" MODE=bench PREC=f64 SDEB200_RNG=2 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/ab_r2y2.txt
# stream v1 (the default stream) with the polar step: cur = v1 without it
VARIANTS="cur polar1" MODE=bench PREC=f64 SDEB200_RNG=0 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/ab_r2y2_v1.txt
SDEB200_LIB=variants/libsdeb200_polar1.so timeout 600 python -m pytest tests -m gpu -x -q -k "parity or stream or generated or fp32" > gpurun_out/pytest_polar1.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_polar1.log; tail -n 5 gpurun_out/pytest_polar1.log
