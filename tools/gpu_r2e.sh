#!/bin/bash
# round 2, fifth GPU call: the new reference-order TMA store kernel (K2u) + FP32 MUFU normals: correctness, then C3 timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stream_v2.py tests/test_golden.py tests/test_gpu_generated.py -m gpu -x -q > gpurun_out/pytest_r2e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2e.log
tail -n 12 gpurun_out/pytest_r2e.log
for r in 0 2; do for p in f64 f32; do
echo "== rng=$r $p"; SDEB200_RNG=$r timeout 300 python tools/prof_sample.py $p 10000000 fullpath 2>&1 | grep " 1 " | awk '{print "   ", $1, $3, $(NF-1), $NF}'
done; done 2>&1 | tee gpurun_out/c3_r2e.log
