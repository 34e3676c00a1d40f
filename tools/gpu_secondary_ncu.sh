#!/bin/bash
# ncu --set full of the secondary kernels in their final form (one profiled launch each; the plain runs come first)
mkdir -p gpurun_out
for m in sample32 simsoa64 mlmc64; do
  case $m in sample32) K=sdek_heston_f_sample_em_f32;; simsoa64) K=sdek_bs_f_simsoa_mil_f64;; mlmc64) K=sdek_heston_f_mlmc_em_f64;; esac
  python tools/prof_kernels.py $m > gpurun_out/sec_${m}_plain.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_r01e_$m python tools/prof_kernels.py $m > gpurun_out/sec_${m}_ncu.log 2>&1
  tail -n 1 gpurun_out/sec_${m}_plain.log
done
ls -la gpurun_out/prof_r01e_*.ncu-rep
