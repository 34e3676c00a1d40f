#!/bin/bash
# end of round 2: the kernels changed after the r02z set (coupled level, K2u FP32 ring, [time][path] addressing) under ncu,
# every configuration once more on stream v2, smoke and the bench line of the final build
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r2zz.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r2zz.log; tail -n 2 gpurun_out/smoke_r2zz.log
timeout 900 python bench.py > gpurun_out/bench_r2zz.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_r2zz.log; tail -n 2 gpurun_out/bench_r2zz.log | cut -c1-300
SDEB200_RNG=2 timeout 1200 python tools/bench_configs.py > gpurun_out/configs_v2_r2zz.log 2>&1; tail -n 1 gpurun_out/configs_v2_r2zz.log | cut -c1-200
export SDEB200_RNG=2
MODES="simsoa32 simref32 mlmc64"
for m in $MODES; do
  case $m in simsoa32) K=sdek_bs_f_simsoa_mil_f32;; simref32) K=sdek_bs_f_simrt_mil_f32;; mlmc64) K=sdek_heston_f_mlmc_em_f64v2;; esac
  timeout 300 python tools/prof_kernels.py $m > gpurun_out/sec2_${m}_plain.log 2>&1 &&
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o /tmp/prof_r02zz_$m python tools/prof_kernels.py $m > gpurun_out/sec2_${m}_ncu.log 2>&1
  tail -n 1 gpurun_out/sec2_${m}_plain.log
done
unset SDEB200_RNG
python tools/ncu_summary.py gpurun_out/r02zz_secondary_ncu_summary.csv $(for m in $MODES; do echo -n "$m=/tmp/prof_r02zz_$m.ncu-rep "; done) | tail -n 1
