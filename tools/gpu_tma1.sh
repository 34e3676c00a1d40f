#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "wiener or multilevel or out_of_bounds or large_host or tma" > gpurun_out/pytest_tma.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tma.log
tail -n 3 gpurun_out/pytest_tma.log
for p in f64 f32; do python tools/prof_sample.py $p 8000000 wiener 2>&1 | grep " 2 wiener"; done
