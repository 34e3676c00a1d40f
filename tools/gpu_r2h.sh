#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 8 gpurun_out/pytest_gpu.log
for r in 0 2; do for p in f64 f32; do
if [ $r = 0 ] && [ $p = f32 ]; then continue; fi
echo "== rng=$r $p"; SDEB200_RNG=$r timeout 300 python tools/prof_sample.py $p 10000000 fullpath 2>&1 | grep " 1 " | awk '{print "   ", $1, $3, $(NF-1), $NF}'
done; done 2>&1 | tee gpurun_out/c3_r2h.log
