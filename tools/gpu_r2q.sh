#!/bin/bash
# sector alignment (32 bytes) for K3t / K7t (variant wt32) and for K2u FP64 (variant a64): parity with each library, then timing
mkdir -p gpurun_out
K="wiener_driven or logtransition or reference_order or layouts_agree or four_dimensional or multilevel_testset or device_resident"
for v in cur wt32 a64; do
  if [ $v = cur ]; then L=""; else L="variants/libsdeb200_$v.so"; fi
  SDEB200_LIB=$L timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_generated.py tests/test_gpu_fp32_parity.py -m gpu -x -q -k "$K" > gpurun_out/pytest_r2q_$v.log 2>&1; echo "pytest $v rc=$?" | tee -a gpurun_out/pytest_r2q_$v.log
  tail -n 3 gpurun_out/pytest_r2q_$v.log | head -n 2
done
export SDEB200_RNG=2
{
for v in cur wt32; do
  if [ $v = cur ]; then L=""; else L="variants/libsdeb200_$v.so"; fi
  for p in f64 f32; do
    echo "== $v wiener $p"; SDEB200_LIB=$L timeout 300 python tools/prof_sample.py $p 8000000 wiener 2>&1 | grep " 2 "
  done
  echo "== $v logt"; SDEB200_LIB=$L LOGT_BOTH=0 timeout 300 python tools/prof_logt.py 2000000 2>&1
done
for v in cur a64; do
  if [ $v = cur ]; then L=""; else L="variants/libsdeb200_$v.so"; fi
  echo "== $v f64 512"; SDEB200_LIB=$L timeout 300 python tools/prof_c3.py f64 10000000 512 reference 2>&1 | tail -n 1
  echo "== $v f64 wiener!"; SDEB200_LIB=$L timeout 300 python tools/prof_sample.py f64 8000000 wiener 2>&1 | grep " 2 " | cut -c1-80
done
} 2>&1 | tee gpurun_out/align_r2q.log
