#!/bin/bash
# Build a tuning variant of the library without touching the in-tree build: tools/build_variant.sh <name> <extra nvcc flags...>
# -> variants/libsdeb200_<name>.so (select it with SDEB200_LIB=...; tools/gpu_ab.sh)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=/tmp/sdeb200_var_$name
rm -rf $tmp && mkdir -p $tmp/sdemodels.jl_b200 $tmp/include
cp -r $root/sdemodels.jl_b200/csrc $tmp/sdemodels.jl_b200/ && rm -rf $tmp/sdemodels.jl_b200/csrc/build
cp $root/include/sdeb200.h $tmp/include/
make -C $tmp/sdemodels.jl_b200/csrc -j8 EXTRA="$*" > $tmp/build.log 2>&1 || { tail -30 $tmp/build.log; exit 1; }
mkdir -p $root/variants && cp $tmp/sdemodels.jl_b200/libsdeb200.so $root/variants/libsdeb200_$name.so
echo "built variants/libsdeb200_$name.so ($*)"
