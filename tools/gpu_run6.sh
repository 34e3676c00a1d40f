#!/bin/bash
mkdir -p gpurun_out
./tools/mb/cost 2>&1 | head -4 > gpurun_out/cost2.log
python tools/prof_sample.py f64 > gpurun_out/v_base.log 2>&1
SDEB200_LIB=$PWD/variants/lib_r7.so python tools/prof_sample.py f64 > gpurun_out/v_r7.log 2>&1
SDEB200_LIB=$PWD/variants/lib_r1.so python tools/prof_sample.py f64 > gpurun_out/v_r1.log 2>&1
for f in cost2 v_base v_r7 v_r1; do echo "== $f"; tail -n 4 gpurun_out/$f.log; done
