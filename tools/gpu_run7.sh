#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 1500 python tools/bench_configs.py > gpurun_out/configs.log 2>&1; echo "rc=$?" >> gpurun_out/configs.log
python tools/prof_sample.py f64 > gpurun_out/prof_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:sdek_heston_f_sample_em_f64 -s 1 -c 1 -f -o gpurun_out/prof_r01b_heston_sample_f64 python tools/prof_sample.py f64 > gpurun_out/ncu_full.log 2>&1
for f in pytest_gpu configs prof_plain; do echo "== $f"; tail -n 12 gpurun_out/$f.log; done
