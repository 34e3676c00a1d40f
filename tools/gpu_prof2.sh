#!/bin/bash
mkdir -p gpurun_out
for spec in "sample32:sdek_heston_f_sample_em_f32" "simsoa64:sdek_bs_f_simsoa_mil_f64" "simsoa32:sdek_bs_f_simsoa_mil_f32" "simw64:sdek_bs_f_simw_em_f64" "mlmc64:sdek_heston_f_mlmc_em_f64"; do
  mode=${spec%%:*}; kern=${spec#*:}
  python tools/prof_kernels.py $mode > gpurun_out/plain_$mode.log 2>&1 &&
  ncu --set full --clock-control none -k regex:$kern -s 1 -c 1 -f -o gpurun_out/prof_r01_$mode python tools/prof_kernels.py $mode > gpurun_out/ncu_$mode.log 2>&1
  tail -1 gpurun_out/plain_$mode.log
done
