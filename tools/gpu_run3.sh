#!/bin/bash
mkdir -p gpurun_out
./tools/mb/mix > gpurun_out/mix.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/prof_sample.py f64 > gpurun_out/ppt2.log 2>&1
SDEB200_LIB=$PWD/variants/libsdeb200_ppt1.so python tools/prof_sample.py f64 > gpurun_out/ppt1.log 2>&1
SDEB200_LIB=$PWD/variants/libsdeb200_ppt4.so python tools/prof_sample.py f64 > gpurun_out/ppt4.log 2>&1
python tools/prof_sample.py f32 > gpurun_out/ppt2_f32.log 2>&1
SDEB200_LIB=$PWD/variants/libsdeb200_ppt4.so python tools/prof_sample.py f32 > gpurun_out/ppt4_f32.log 2>&1
for f in mix pytest_gpu ppt1 ppt2 ppt4 ppt2_f32 ppt4_f32; do echo "== $f"; tail -n 4 gpurun_out/$f.log; done
