#!/bin/bash
# 2 paths per thread in 256-thread CTAs + unconditional main loop: whole GPU suite, A/B against 4 paths per thread, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2y.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_r2y.log
tail -n 8 gpurun_out/pytest_r2y.log
VARIANTS="cur v2ppt4" MODE=bench PREC=f64 SDEB200_RNG=2 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/ab_r2y.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2>&1 | tail -n 1 | tee gpurun_out/bench_r2y.json | cut -c1-300
