#!/bin/bash
# what the driver runs at round end, in its order, plus every BASELINE configuration (tools/bench_configs.py)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -n 3 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1; echo "reference rc=$?" >> gpurun_out/bench_reference.log; tail -n 2 gpurun_out/bench_reference.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_full.log; tail -n 3 gpurun_out/bench_full.log
if [ "$1" = "configs" ]; then
timeout 1500 python tools/bench_configs.py > gpurun_out/configs.log 2>&1; echo "rc=$?" >> gpurun_out/configs.log; tail -n 30 gpurun_out/configs.log
fi
