#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/prof_sample.py f64 8000000 wiener > gpurun_out/w64.log 2>&1
python tools/prof_sample.py f32 8000000 wiener > gpurun_out/w32.log 2>&1
python tools/prof_sample.py f64 4000000 simulate > gpurun_out/sim64.log 2>&1
python tools/prof_sample.py f32 4000000 simulate > gpurun_out/sim32.log 2>&1
for f in pytest_gpu w64 w32 sim64 sim32; do echo "== $f"; tail -n 5 gpurun_out/$f.log; done
