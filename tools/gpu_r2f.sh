#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stream_v2.py tests/test_golden.py tests/test_gpu_generated.py -m gpu -q > gpurun_out/pytest_r2f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2f.log
tail -n 12 gpurun_out/pytest_r2f.log
SDEB200_RNG=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sdek_bs_f_simrt_mil_f(32|64v2)" -c 4 -f -o gpurun_out/prof_r02b_simrt python tools/prof_simrt.py > gpurun_out/ncu_r2f.log 2>&1
tail -n 3 gpurun_out/ncu_r2f.log
