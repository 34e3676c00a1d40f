#!/bin/bash
# final code of the round (polar step on both streams + in generated functors, unconditional v2 loop): GPU suite first;
# then A/B of 2 against 4 paths per thread on stream v2, v1 timing, generated-vs-built-in timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2y3.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_r2y3.log
tail -n 8 gpurun_out/pytest_r2y3.log
VARIANTS="cur p2" MODE=bench PREC=f64 SDEB200_RNG=2 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/ab_r2y3.txt
VARIANTS="cur" MODE=bench PREC=f64 SDEB200_RNG=0 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/ab_r2y3_v1.txt
for r in 0 2; do echo "== generated vs built-in, SDEB200_RNG=$r"; SDEB200_RNG=$r timeout 600 python tools/prof_generated.py 2>&1 | tail -n 6; done | tee gpurun_out/generated_r2y3.txt
