#!/bin/bash
# polar form of the Heston step (one square root of V*w): whole GPU suite, then headline + C4 timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2v.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_r2v.log
tail -n 15 gpurun_out/pytest_r2v.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2>&1 | tail -n 1 | tee gpurun_out/bench_r2v.json | cut -c1-600
SDEB200_RNG=2 timeout 200 python tools/mlmc_timing.py 2>&1 | tail -n 2 | tee gpurun_out/mlmc_r2v.log
