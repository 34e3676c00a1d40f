#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_generated.py tests/test_gpu_stream_v2.py -m gpu -x -q -k "reference_order or layouts_agree or four_dimensional or multilevel_testset or kernel_forms or two_states" > gpurun_out/pytest_r2s.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_r2s.log
tail -n 4 gpurun_out/pytest_r2s.log
export SDEB200_RNG=2
{
for p in f64 f32; do
  echo "== $p 512"; timeout 300 python tools/prof_c3.py $p 10000000 512 reference 2>&1 | tail -n 1
  timeout 300 python tools/prof_sample.py $p 8000000 wiener 2>&1 | grep " 2 " | cut -c1-70
done
timeout 300 python tools/prof_c3.py f64 10000000 511 reference 2>&1 | tail -n 1
SDEB200_RNG=0 timeout 300 python tools/prof_c3.py f64 10000000 512 reference 2>&1 | tail -n 1
} 2>&1 | tee gpurun_out/c3_r2s.log
