#!/bin/bash
# ncu --set full capture of the headline kernel in its bench shape (terminal states kept); the plain run comes first
mkdir -p gpurun_out
python tools/prof_sample.py f64 4849664 bench > gpurun_out/prof_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:sdek_heston_f_sample_em_f64 -s 1 -c 1 -f -o gpurun_out/prof_r01f_heston_sample_f64 python tools/prof_sample.py f64 4849664 bench > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/prof_plain.log; tail -n 2 gpurun_out/ncu_full.log
