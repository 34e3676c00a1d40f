#!/bin/bash
# ncu --set full of the kernels changed in this session: K7t (logtransition), K2u FP32 (64-byte sector-aligned tiles), K2u FP64
mkdir -p gpurun_out
export SDEB200_RNG=2
for m in logt64 simref32 simref64; do
  case $m in logt64) K=sdek_bs_f_logtt_f64;; simref32) K=sdek_bs_f_simrt_mil_f32;; simref64) K=sdek_bs_f_simrt_mil_f64v2;; esac
  timeout 300 python tools/prof_kernels.py $m > gpurun_out/r2t_${m}_plain.log 2>&1 &&
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o /tmp/prof_r02t_$m python tools/prof_kernels.py $m > gpurun_out/r2t_${m}_ncu.log 2>&1
  tail -n 1 gpurun_out/r2t_${m}_plain.log
done
python tools/ncu_summary.py gpurun_out/r02t_changed_kernels_ncu_summary.csv $(for m in logt64 simref32 simref64; do echo -n "$m=/tmp/prof_r02t_$m.ncu-rep "; done) | tail -n 1
for m in logt64 simref32; do ncu -i /tmp/prof_r02t_$m.ncu-rep --page details --csv > gpurun_out/r02t_${m}_details.csv 2>/dev/null; done
ls -la /tmp/*.ncu-rep | awk '{print $5, $9}'
