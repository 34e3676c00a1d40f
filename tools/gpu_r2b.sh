#!/bin/bash
# round 2, second GPU call: FP32 parity + v2 micro-variants + ncu of the v2 headline kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fp32_parity.py tests/test_gpu_stream_v2.py "tests/test_gpu_parity.py::test_wiener_driven_f32_within_1e5" -m gpu -x -q > gpurun_out/pytest_r2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2b.log
tail -n 15 gpurun_out/pytest_r2b.log
for round in 1 2; do
for v in cur imad rot1 both ph7; do
if [ $v = cur ]; then L=""; else L="variants/libsdeb200_$v.so"; fi
for r in 0 2; do
if [ $r = 0 ] && [ $v != cur ] && [ $v != ph7 ]; then continue; fi
echo -n "$v rng=$r: "; SDEB200_RNG=$r SDEB200_LIB=$L timeout 300 python tools/prof_sample.py f64 19398656 bench 2>&1 | tail -n 1 | awk '{print $(NF-3), $(NF-2), $(NF-1)}'
done; done; done 2>&1 | tee gpurun_out/ab_r2b.log
SDEB200_RNG=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sdek_heston_f_sample_em_f64v2 -s 1 -c 1 -f -o gpurun_out/prof_r02a_heston_sample_f64v2 python tools/prof_sample.py f64 4849664 bench > gpurun_out/ncu_r2b.log 2>&1
tail -n 2 gpurun_out/ncu_r2b.log
