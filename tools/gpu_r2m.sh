#!/bin/bash
# round 2 (second session): K7 fast math + K7t parity and timing; C3 FP32 reference-order diagnostics
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_generated.py -m gpu -x -q -k "logtransition" > gpurun_out/pytest_r2m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2m.log
tail -n 5 gpurun_out/pytest_r2m.log
timeout 600 python tools/prof_logt.py 2000000 2>&1 | tee gpurun_out/logt_r2m.log
export SDEB200_RNG=2
{
for v in cur diag1 diag2; do
  if [ $v = cur ]; then L=""; else L="variants/libsdeb200_$v.so"; fi
  for p in f32 f64; do
    echo "== $v $p 512"; SDEB200_LIB=$L timeout 300 python tools/prof_c3.py $p 10000000 512 reference 2>&1 | tail -n 1
  done
done
for ns in 511 1023 2047 127; do
  echo "== cur f32 $ns"; timeout 300 python tools/prof_c3.py f32 $((10000000*512/(ns+1))) $ns reference,soa 2>&1 | tail -n 2
  echo "== cur f64 $ns"; timeout 300 python tools/prof_c3.py f64 $((10000000*512/(ns+1))) $ns reference 2>&1 | tail -n 1
done
} 2>&1 | tee gpurun_out/c3_r2m.log
