#!/bin/bash
# A/B of library variants (variants/libsdeb200_<name>.so, "cur" = the in-tree build): VARIANTS="a b" MODE=bench|paths PREC="f64 f32"
MODE=${MODE:-bench}; PREC=${PREC:-f64}
for round in 1 2; do
for v in $VARIANTS; do
if [ $v = cur ]; then L=""; else L="variants/libsdeb200_$v.so"; fi
for p in $PREC; do
if [ $MODE = bench ]; then
echo -n "$v $p: "; SDEB200_LIB=$L python tools/prof_sample.py $p 19398656 bench 2>&1 | tail -n 1 | awk '{print $(NF-3), $(NF-2), $(NF-1)}'
else
echo "$v $p:"; SDEB200_LIB=$L python tools/prof_sample.py $p 4000000 paths 2>&1 | grep " 1 " | awk '{print "   ", $1, $(NF-1), $NF}'
fi
done; done; done
