#!/bin/bash
mkdir -p gpurun_out
timeout 600 tools/mb/simw_tma 2000000 513 dbg > gpurun_out/mb_simw_tma.txt 2>&1; echo "rc=$?" >> gpurun_out/mb_simw_tma.txt
cat gpurun_out/mb_simw_tma.txt
