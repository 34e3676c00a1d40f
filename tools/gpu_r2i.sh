#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity_scale.py tests/test_gpu_multidevice.py -m gpu -q --durations=6 > gpurun_out/pytest_r2i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2i.log
tail -n 14 gpurun_out/pytest_r2i.log
