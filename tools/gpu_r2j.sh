#!/bin/bash
# round 2: final measurements — bench (all legs), every configuration on both streams, ncu captures, launch list
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2j.log 2>&1; tail -n 1 gpurun_out/bench_r2j.log | cut -c1-400
SDEB200_RNG=2 timeout 1200 python tools/bench_configs.py > gpurun_out/configs_v2.log 2>&1; tail -n 1 gpurun_out/configs_v2.log
SDEB200_RNG=0 timeout 1200 python tools/bench_configs.py > gpurun_out/configs_v1.log 2>&1; tail -n 1 gpurun_out/configs_v1.log
# headline kernel at the bench shape (10^8 paths, terminal states kept): one launch under ncu --set full
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sdek_heston_f_sample_em_f64v2 -s 3 -c 1 -f -o /tmp/prof_r02c_headline python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_headline.log 2>&1; tail -n 1 gpurun_out/ncu_headline.log | cut -c1-200
export SDEB200_RNG=2
for m in sample32 simsoa64 simsoa32 simref64 simref32 mlmc64 simw64 logt64; do
  case $m in sample32) K=sdek_heston_f_sample_em_f32;; simsoa64) K=sdek_bs_f_simsoa_mil_f64v2;; simsoa32) K=sdek_bs_f_simsoa_mil_f32;; simref64) K=sdek_bs_f_simrt_mil_f64v2;; simref32) K=sdek_bs_f_simrt_mil_f32;; mlmc64) K=sdek_heston_f_mlmc_em_f64v2;; simw64) K=sdek_bs_f_simwt_em_f64;; logt64) K=sdek_bs_f_logt_f64;; esac
  timeout 300 python tools/prof_kernels.py $m > gpurun_out/sec_${m}_plain.log 2>&1 &&
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o /tmp/prof_r02c_$m python tools/prof_kernels.py $m > gpurun_out/sec_${m}_ncu.log 2>&1
  tail -n 1 gpurun_out/sec_${m}_plain.log
done
unset SDEB200_RNG
# condense the captures here (the reports with sources are too large to travel back); keep the headline report itself
python tools/ncu_summary.py gpurun_out/r02c_headline_ncu_summary.csv headline=/tmp/prof_r02c_headline.ncu-rep | tail -n 1
python tools/ncu_summary.py gpurun_out/r02c_secondary_ncu_summary.csv $(for m in sample32 simsoa64 simsoa32 simref64 simref32 mlmc64 simw64 logt64; do echo -n "$m=/tmp/prof_r02c_$m.ncu-rep "; done) | tail -n 1
ncu -i /tmp/prof_r02c_headline.ncu-rep --page source --csv > gpurun_out/r02c_headline_source.csv 2>/dev/null; ls -la /tmp/*.ncu-rep | awk '{print $5, $9}'
timeout 600 python bench.py --steps 2 --warmup 3 --paths 20000000 --no-cpu-baseline > gpurun_out/bench_small.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02c.csv python bench.py --steps 2 --warmup 3 --paths 20000000 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
grep -c sdek gpurun_out/launches_r02c.csv
