#!/bin/bash
# launch list of the bench command (per-launch device times, cold-cache and serialised: compare shares, not absolutes)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --paths 20000000 --no-cpu-baseline > gpurun_out/bench_small.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 2 --warmup 3 --paths 20000000 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
tail -n 1 gpurun_out/bench_small.log | cut -c1-200; grep -c sdek gpurun_out/launches_r01d.csv
