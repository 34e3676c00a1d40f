#!/bin/bash
# per-instruction stall sampling of the reference-order FP32 trajectory kernel (K2u) and the [time][path] FP32 kernel
mkdir -p gpurun_out
export SDEB200_RNG=2
for m in simref32 simsoa32; do
  case $m in simref32) K=sdek_bs_f_simrt_mil_f32;; simsoa32) K=sdek_bs_f_simsoa_mil_f32;; esac
  timeout 300 python tools/prof_kernels.py $m > gpurun_out/src_${m}_plain.log 2>&1 &&
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o /tmp/prof_src_$m python tools/prof_kernels.py $m > gpurun_out/src_${m}_ncu.log 2>&1
  ncu -i /tmp/prof_src_$m.ncu-rep --page source --csv > gpurun_out/r02zz_${m}_source.csv 2>/dev/null
  ls -la gpurun_out/r02zz_${m}_source.csv
done
