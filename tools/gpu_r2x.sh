#!/bin/bash
# occupancy variants of the polar headline kernel; FP32 angle-from-mantissa experiment on the C3 kernels
mkdir -p gpurun_out
VARIANTS="cur ppt2 ppt2mb5 ppt2mb6 ppt1mb8 ppt1mb6" MODE=bench PREC=f64 SDEB200_RNG=2 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/ab_r2x.txt
export SDEB200_RNG=2
{
for round in 1 2; do
for v in cur f32ang; do
  if [ $v = cur ]; then L=""; else L="variants/libsdeb200_$v.so"; fi
  echo "== $v"; SDEB200_LIB=$L timeout 300 python tools/prof_c3.py f32 10000000 512 reference,soa 2>&1 | tail -n 2
  echo -n "$v f32 bench: "; SDEB200_LIB=$L python tools/prof_sample.py f32 19398656 bench 2>&1 | tail -n 1 | awk '{print $(NF-3), $(NF-2), $(NF-1)}'
done; done
} 2>&1 | tee gpurun_out/f32ang_r2x.txt
