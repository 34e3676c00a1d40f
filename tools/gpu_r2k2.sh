#!/bin/bash
# per-instruction stall sampling of the shipped headline kernel (bench-shaped launch, 19.4 M paths) and of the coupled level kernel
mkdir -p gpurun_out
export SDEB200_RNG=2
timeout 300 python tools/prof_sample.py f64 19398656 bench > gpurun_out/src_headline_plain.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sdek_heston_f_sample_em_f64v2 -s 1 -c 1 -f -o /tmp/prof_src_headline python tools/prof_sample.py f64 19398656 bench > gpurun_out/src_headline_ncu.log 2>&1
ncu -i /tmp/prof_src_headline.ncu-rep --page source --csv > gpurun_out/r02zz_headline_source.csv 2>/dev/null
timeout 300 python tools/prof_kernels.py mlmc64 > gpurun_out/src_mlmc_plain.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sdek_heston_f_mlmc_em_f64v2 -s 1 -c 1 -f -o /tmp/prof_src_mlmc python tools/prof_kernels.py mlmc64 > gpurun_out/src_mlmc_ncu.log 2>&1
ncu -i /tmp/prof_src_mlmc.ncu-rep --page source --csv > gpurun_out/r02zz_mlmc64_source.csv 2>/dev/null
ls -la gpurun_out/r02zz_headline_source.csv gpurun_out/r02zz_mlmc64_source.csv; tail -n 1 gpurun_out/src_headline_plain.log
