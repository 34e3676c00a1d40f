#!/bin/bash
# ring depth of the reference-order TMA store kernel (K2u) after the 64-byte FP32 tiles: stages x precision; wiener! as well
mkdir -p gpurun_out
export SDEB200_RNG=2
{
for round in 1 2; do
for v in cur s5 s6 s8; do
  if [ $v = cur ]; then L=""; else L="variants/libsdeb200_$v.so"; fi
  echo -n "$v: "; SDEB200_LIB=$L timeout 300 python tools/prof_c3.py f32 10000000 512 reference 2>&1 | tail -n 1
done
for v in cur s4f64 s6f64; do
  if [ $v = cur ]; then L=""; else L="variants/libsdeb200_$v.so"; fi
  echo -n "$v: "; SDEB200_LIB=$L timeout 300 python tools/prof_c3.py f64 10000000 512 reference 2>&1 | tail -n 1
done
done
for v in cur s6 s8; do
  if [ $v = cur ]; then L=""; else L="variants/libsdeb200_$v.so"; fi
  echo "$v wiener f32:"; SDEB200_LIB=$L timeout 300 python tools/prof_sample.py f32 8000000 wiener 2>&1 | grep " 2 " | cut -c1-90
done
for v in cur s4f64 s6f64; do
  if [ $v = cur ]; then L=""; else L="variants/libsdeb200_$v.so"; fi
  echo "$v wiener f64:"; SDEB200_LIB=$L timeout 300 python tools/prof_sample.py f64 8000000 wiener 2>&1 | grep " 2 " | cut -c1-90
done
} 2>&1 | tee gpurun_out/r02zz_c3_stages2.txt
