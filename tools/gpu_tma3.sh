#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tma or layouts or oracle_rows or out_of_bounds or multilevel or exact" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
for p in f64 f32; do
python tools/prof_sample.py $p 4000000 paths 2>&1 | tail -n 1
python tools/prof_sample.py $p 8000000 wiener 2>&1 | grep " 2 wiener"
done
