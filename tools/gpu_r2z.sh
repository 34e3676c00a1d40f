#!/bin/bash
# round 2, final measurement set (polar Heston step, unconditional v2 time loop): the driver's round-end sequence, every
# configuration on both streams, ncu captures, launch list.  Files land in gpurun_out/ with the prefix r02z.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r2z.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_r2z.log; tail -n 3 gpurun_out/pytest_r2z.log
if [ $rc -ne 0 ]; then tail -n 60 gpurun_out/pytest_r2z.log; exit 1; fi
# FP32 terminal kernel: polar step (cur) against the build before it (prev)
VARIANTS="cur prev" MODE=bench PREC=f32 SDEB200_RNG=2 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02z_ab_f32_polar.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r2z.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r2z.log; tail -n 2 gpurun_out/smoke_r2z.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference_r2z.log 2>&1; echo "reference rc=$?" >> gpurun_out/bench_reference_r2z.log; tail -n 2 gpurun_out/bench_reference_r2z.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_r2z.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_r2z.log; tail -n 2 gpurun_out/bench_r2z.log | cut -c1-400
SDEB200_RNG=2 timeout 1200 python tools/bench_configs.py > gpurun_out/configs_v2_r2z.log 2>&1; tail -n 1 gpurun_out/configs_v2_r2z.log | cut -c1-300
SDEB200_RNG=0 timeout 1200 python tools/bench_configs.py > gpurun_out/configs_v1_r2z.log 2>&1; tail -n 1 gpurun_out/configs_v1_r2z.log | cut -c1-300
# headline kernel at the bench shape (10^8 paths, terminal states kept): one launch under ncu --set full
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sdek_heston_f_sample_em_f64v2 -s 3 -c 1 -f -o /tmp/prof_r02z_headline python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_headline_r2z.log 2>&1; tail -n 1 gpurun_out/ncu_headline_r2z.log | cut -c1-200
export SDEB200_RNG=2
MODES="sample32 simsoa64 simsoa32 simref64 simref32 mlmc64 simw64 logt64"
for m in $MODES; do
  case $m in sample32) K=sdek_heston_f_sample_em_f32;; simsoa64) K=sdek_bs_f_simsoa_mil_f64v2;; simsoa32) K=sdek_bs_f_simsoa_mil_f32;; simref64) K=sdek_bs_f_simrt_mil_f64v2;; simref32) K=sdek_bs_f_simrt_mil_f32;; mlmc64) K=sdek_heston_f_mlmc_em_f64v2;; simw64) K=sdek_bs_f_simwt_em_f64;; logt64) K=sdek_bs_f_logtt_f64;; esac
  timeout 300 python tools/prof_kernels.py $m > gpurun_out/sec_${m}_plain.log 2>&1 &&
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o /tmp/prof_r02z_$m python tools/prof_kernels.py $m > gpurun_out/sec_${m}_ncu.log 2>&1
  tail -n 1 gpurun_out/sec_${m}_plain.log
done
unset SDEB200_RNG
python tools/ncu_summary.py gpurun_out/r02z_headline_ncu_summary.csv headline=/tmp/prof_r02z_headline.ncu-rep | tail -n 1
python tools/ncu_summary.py gpurun_out/r02z_secondary_ncu_summary.csv $(for m in $MODES; do echo -n "$m=/tmp/prof_r02z_$m.ncu-rep "; done) | tail -n 1
timeout 600 python bench.py --steps 2 --warmup 3 --paths 20000000 --no-cpu-baseline > gpurun_out/bench_small_r2z.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches_bench_small.csv python bench.py --steps 2 --warmup 3 --paths 20000000 --no-cpu-baseline > gpurun_out/ncu_list_r2z.log 2>&1
grep -c sdek gpurun_out/r02z_launches_bench_small.csv
