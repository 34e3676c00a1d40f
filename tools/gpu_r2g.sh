#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stream_v2.py tests/test_golden.py tests/test_gpu_generated.py -m gpu -x -q > gpurun_out/pytest_r2g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2g.log
tail -n 8 gpurun_out/pytest_r2g.log
for v in cur st2 st4; do
if [ $v = cur ]; then L=""; else L="variants/libsdeb200_$v.so"; fi
for r in 0 2; do for p in f64 f32; do
if [ $r = 0 ] && [ $p = f32 ]; then continue; fi
echo "== $v rng=$r $p"; SDEB200_LIB=$L SDEB200_RNG=$r timeout 300 python tools/prof_sample.py $p 10000000 fullpath 2>&1 | grep " 1 " | awk '{print "   ", $1, $3, $(NF-1), $NF}'
done; done; done 2>&1 | tee gpurun_out/c3_r2g.log
