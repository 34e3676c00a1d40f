#!/bin/bash
# round 2, fourth GPU call: whole GPU suite, bench with all legs, every configuration on both normal streams
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 25 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2d.log 2>&1; tail -n 1 gpurun_out/bench_r2d.log | cut -c1-1500
SDEB200_RNG=2 timeout 1200 python tools/bench_configs.py > gpurun_out/configs_v2.log 2>&1; tail -n 3 gpurun_out/configs_v2.log | cut -c1-600
SDEB200_RNG=0 timeout 1200 python tools/bench_configs.py > gpurun_out/configs_v1.log 2>&1; tail -n 1 gpurun_out/configs_v1.log
