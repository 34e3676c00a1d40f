#!/bin/bash
# coarse step of a coupled level from the two polar triples (Heston::step_polar2): GPU suite on the variant, C4 timing A/B
mkdir -p gpurun_out
export SDEB200_RNG=2
for r in 1 2; do
echo -n "cur: "; timeout 200 python tools/mlmc_timing.py 2>&1 | tail -n 1
echo -n "polar2c: "; SDEB200_LIB=$PWD/variants/libsdeb200_polar2c.so timeout 200 python tools/mlmc_timing.py 2>&1 | tail -n 1
done | tee gpurun_out/mlmc_r2h.txt
unset SDEB200_RNG
export SDEB200_LIB=$PWD/variants/libsdeb200_polar2c.so
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2h.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_r2h.log
tail -n 8 gpurun_out/pytest_r2h.log
