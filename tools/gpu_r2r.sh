#!/bin/bash
# K2u with sector-aligned tiles in both precisions (straddling batches): the whole GPU suite, then C3 / wiener! timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2r.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_r2r.log
tail -n 4 gpurun_out/pytest_r2r.log
export SDEB200_RNG=2
{
for p in f64 f32; do
  echo "== $p 512"; timeout 300 python tools/prof_c3.py $p 10000000 512 reference,soa 2>&1 | tail -n 2
  timeout 300 python tools/prof_sample.py $p 8000000 wiener 2>&1 | grep " 2 "
done
SDEB200_RNG=0 timeout 300 python tools/prof_c3.py f64 10000000 512 reference,soa 2>&1 | tail -n 2
} 2>&1 | tee gpurun_out/c3_r2r.log
