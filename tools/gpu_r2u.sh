#!/bin/bash
# after the loop-overhead trims (K2u acquire per tile / uniform warp index, K7t result store loop): whole GPU suite, then timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2u.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_r2u.log
tail -n 4 gpurun_out/pytest_r2u.log
export SDEB200_RNG=2
{
for p in f64 f32; do
  echo "== $p 512"; timeout 300 python tools/prof_c3.py $p 10000000 512 reference,soa 2>&1 | tail -n 2
  timeout 300 python tools/prof_sample.py $p 8000000 wiener 2>&1 | grep " 2 "
done
LOGT_BOTH=0 timeout 300 python tools/prof_logt.py 2000000
} 2>&1 | tee gpurun_out/c3_r2u.log
