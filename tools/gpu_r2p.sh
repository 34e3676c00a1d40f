#!/bin/bash
# K2u FP32 with 64-byte tile rows aligned to 32-byte sectors (g <= 8): parity and C3 timing per ring depth
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_generated.py tests/test_gpu_stream_v2.py -m gpu -x -q -k "reference_order or layouts_agree or multilevel_testset or kernel_forms or four_dimensional" > gpurun_out/pytest_r2p.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2p.log
tail -n 5 gpurun_out/pytest_r2p.log
export SDEB200_RNG=2
{
for v in cur s3 s6 r128; do
  if [ $v = cur ]; then L=""; else L="variants/libsdeb200_$v.so"; fi
  echo "== $v f32 512"; SDEB200_LIB=$L timeout 300 python tools/prof_c3.py f32 10000000 512 reference 2>&1 | tail -n 1
done
echo "== cur f32 511/1023/other"; timeout 300 python tools/prof_c3.py f32 10000000 511 reference 2>&1 | tail -n 1; timeout 300 python tools/prof_c3.py f32 5000000 1023 reference 2>&1 | tail -n 1
timeout 300 python tools/prof_c3.py f32 10000000 513 reference 2>&1 | tail -n 1; timeout 300 python tools/prof_c3.py f32 10000000 514 reference 2>&1 | tail -n 1; timeout 300 python tools/prof_c3.py f32 10000000 515 reference 2>&1 | tail -n 1
} 2>&1 | tee gpurun_out/c3_r2p.log
timeout 300 python tools/prof_sample.py f32 8000000 wiener 2>&1 | grep " 2 " | tee gpurun_out/wiener_r2p.log
