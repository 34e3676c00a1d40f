#!/bin/bash
# FP32 ring of six stages in the tree: GPU suite; FP64 with 64-byte tiles x 4 stages against the shipped 128-byte x 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2m2.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_r2m2.log
tail -n 6 gpurun_out/pytest_r2m2.log
export SDEB200_RNG=2
{
for round in 1 2; do
for v in cur f64r64s4; do
  if [ $v = cur ]; then L=""; else L="variants/libsdeb200_$v.so"; fi
  echo -n "$v: "; SDEB200_LIB=$L timeout 300 python tools/prof_c3.py f64 10000000 512 reference 2>&1 | tail -n 1
done; done
echo -n "cur: "; timeout 300 python tools/prof_c3.py f32 10000000 512 reference,soa 2>&1 | tail -n 2
} 2>&1 | tee gpurun_out/r02zz_c3_f64_tile64.txt
