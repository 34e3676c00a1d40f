#!/bin/bash
mkdir -p gpurun_out
./tools/mb/cost > gpurun_out/cost.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/prof_sample.py f64 > gpurun_out/s64.log 2>&1
for f in cost pytest_gpu s64; do echo "== $f"; tail -n 20 gpurun_out/$f.log; done
