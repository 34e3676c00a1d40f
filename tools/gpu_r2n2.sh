#!/bin/bash
# [time][path] kernel with bumped row pointers and the vector/scalar choice outside the loop: GPU suite, C3 both layouts
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2n2.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_r2n2.log
tail -n 6 gpurun_out/pytest_r2n2.log
export SDEB200_RNG=2
{
for p in f32 f64; do timeout 300 python tools/prof_c3.py $p 10000000 512 reference,soa 2>&1 | tail -n 2; done
SDEB200_RNG=0 timeout 300 python tools/prof_c3.py f64 10000000 512 soa 2>&1 | tail -n 1
} 2>&1 | tee gpurun_out/r02zz_c3_soa_pointer_bump.txt
