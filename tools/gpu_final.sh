#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_full.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?" >> gpurun_out/bench_ref.log
timeout 1500 python tools/bench_configs.py --big > gpurun_out/configs.log 2>&1; echo "rc=$?" >> gpurun_out/configs.log
python tools/prof_sample.py f64 8000000 wiener > gpurun_out/w64.log 2>&1
for f in pytest_gpu smoke bench_full bench_ref w64; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-600; done
