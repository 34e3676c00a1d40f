#!/bin/bash
# A/B of library variants (variants/libsdeb200_<name>.so, "cur" = the in-tree build) on the headline kernel shape
for round in 1 2; do
for v in $VARIANTS; do
if [ $v = cur ]; then L=""; else L="variants/libsdeb200_$v.so"; fi
echo -n "$v f64: "; SDEB200_LIB=$L python tools/prof_sample.py f64 19398656 bench 2>&1 | tail -n 1 | awk '{print $(NF-3), $(NF-2), $(NF-1)}'
done; done
