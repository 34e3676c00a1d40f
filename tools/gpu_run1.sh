#!/bin/bash
# first GPU contact: parity tests, smoke, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 2 --warmup 3 --paths 20000000 > gpurun_out/bench_small.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_small.log
tail -5 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench_small.log
