#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_full.log
python tools/prof_sample.py f64 4849664 bench > gpurun_out/prof_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:sdek_heston_f_sample_em_f64 -s 1 -c 1 -f -o gpurun_out/prof_r01c_heston_sample_f64 python tools/prof_sample.py f64 4849664 bench > gpurun_out/ncu_full.log 2>&1
python bench.py --steps 2 --warmup 3 --paths 20000000 --no-cpu-baseline > gpurun_out/bench_small.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 2 --warmup 3 --paths 20000000 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
timeout 1500 python tools/bench_configs.py > gpurun_out/configs.log 2>&1; echo "rc=$?" >> gpurun_out/configs.log
python tools/prof_sample.py f64 8000000 wiener > gpurun_out/w64.log 2>&1
python tools/prof_sample.py f32 8000000 wiener > gpurun_out/w32.log 2>&1
for f in pytest_gpu bench_full prof_plain w64 w32; do echo "== $f"; tail -n 4 gpurun_out/$f.log; done
