#!/bin/bash
# round 2, first GPU call: stream v2 correctness + A/B of the headline kernel variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream_v2.py tests/test_golden.py -m gpu -x -q > gpurun_out/pytest_v2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v2.log
tail -n 15 gpurun_out/pytest_v2.log
for round in 1 2; do
for v in cur nofine sq7 mb4 mb2; do
if [ $v = cur ]; then L=""; else L="variants/libsdeb200_$v.so"; fi
for r in 0 2; do
echo -n "$v rng=$r: "; SDEB200_RNG=$r SDEB200_LIB=$L timeout 300 python tools/prof_sample.py f64 19398656 bench 2>&1 | tail -n 1 | awk '{print $(NF-3), $(NF-2), $(NF-1)}'
done; done; done 2>&1 | tee gpurun_out/ab_r2a.log
