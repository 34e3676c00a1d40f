#!/bin/bash
mkdir -p gpurun_out
SDEB200_LIB=variants/libsdeb200_man64.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "reference_order or layouts_agree" > gpurun_out/pytest_r2k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2k.log
tail -n 4 gpurun_out/pytest_r2k.log
for v in cur man32 man64; do
if [ $v = cur ]; then L=""; else L="variants/libsdeb200_$v.so"; fi
for r in 0 2; do for p in f64 f32; do
if [ $r = 0 ] && [ $p = f32 ]; then continue; fi
if [ $v = man32 ] && [ $p = f64 ]; then continue; fi
echo "== $v rng=$r $p"; SDEB200_LIB=$L SDEB200_RNG=$r timeout 300 python tools/prof_sample.py $p 10000000 fullpath 2>&1 | grep " 1 " | awk '{print "   ", $1, $3, $(NF-1), $NF}'
done; done; done 2>&1 | tee gpurun_out/c3_r2k.log
