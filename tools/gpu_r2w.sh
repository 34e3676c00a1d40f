#!/bin/bash
# polar headline kernel: A/B of occupancy variants, then one ncu --set full capture at the bench shape
mkdir -p gpurun_out
VARIANTS="cur mb3 ppt2" MODE=bench PREC=f64 SDEB200_RNG=2 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/ab_r2w.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sdek_heston_f_sample_em_f64v2 -s 3 -c 1 -f -o /tmp/prof_r02w_headline python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_headline_r2w.log 2>&1; tail -n 1 gpurun_out/ncu_headline_r2w.log | cut -c1-200
python tools/ncu_summary.py gpurun_out/r02w_headline_ncu_summary.csv headline=/tmp/prof_r02w_headline.ncu-rep | tail -n 1
ncu -i /tmp/prof_r02w_headline.ncu-rep --page source --csv > gpurun_out/r02w_headline_source.csv 2>/dev/null; ls -la /tmp/*.ncu-rep | awk '{print $5, $9}'
