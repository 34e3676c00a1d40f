#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "wiener or multilevel or out_of_bounds or large_host" > gpurun_out/pytest_tma.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tma.log
tail -n 25 gpurun_out/pytest_tma.log
python tools/prof_sample.py f64 8000000 wiener > gpurun_out/w64.log 2>&1; tail -n 6 gpurun_out/w64.log
python tools/prof_sample.py f32 8000000 wiener > gpurun_out/w32.log 2>&1; tail -n 6 gpurun_out/w32.log
SDEB200_NO_TMA=1 python tools/prof_sample.py f64 8000000 wiener > gpurun_out/w64_old.log 2>&1; tail -n 6 gpurun_out/w64_old.log
SDEB200_NO_TMA=1 python tools/prof_sample.py f32 8000000 wiener > gpurun_out/w32_old.log 2>&1; tail -n 6 gpurun_out/w32_old.log
