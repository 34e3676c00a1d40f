#!/bin/bash
mkdir -p gpurun_out
VARIANTS="cur u2" MODE=bench PREC=f64 SDEB200_RNG=2 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/ab_r2zz_unroll.txt
