#!/bin/bash
mkdir -p gpurun_out
for v in st4 st8; do
for p in f64 f32; do
echo "== variant $v $p"; SDEB200_LIB=variants/libsdeb200_$v.so python tools/prof_sample.py $p 8000000 wiener 2>&1 | grep " 2 wiener"
done; done
python tools/prof_sample.py f64 2000000 wiener > gpurun_out/prof_w_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"simwt|wiener1" -s 4 -c 2 -f -o gpurun_out/prof_r01d_simwt_f64 python tools/prof_sample.py f64 2000000 wiener > gpurun_out/ncu_w.log 2>&1
tail -n 3 gpurun_out/ncu_w.log
python tools/prof_sample.py f64 2000000 paths > gpurun_out/prof_p_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"simrt" -s 1 -c 1 -f -o gpurun_out/prof_r01d_simrt_f64 python tools/prof_sample.py f64 2000000 paths > gpurun_out/ncu_p.log 2>&1
tail -n 3 gpurun_out/ncu_p.log
ls -la gpurun_out/*.ncu-rep
