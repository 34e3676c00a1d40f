#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_generated.py tests/test_gpu_parity.py -m gpu -x -q -k "logtransition or device_resident" > gpurun_out/pytest_r2n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2n.log
tail -n 5 gpurun_out/pytest_r2n.log
timeout 600 python tools/prof_logt.py 2000000 2>&1 | tee gpurun_out/logt_r2n.log
