#!/bin/bash
# round 2, final code on one 8 x B200 box: bit-identity of the sharded sums, bench.py at N = 8 (weak, strong, e2e legs)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02z_gpus8.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 tools/dist_check.py > gpurun_out/r02z_dist_check8.log 2>&1; echo "rc=$?" >> gpurun_out/r02z_dist_check8.log
grep -h "rank\|rc=" gpurun_out/r02z_dist_check8.log | cut -c1-170 | tail -n 3
timeout 900 $TR --master-port 29523 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02z_bench_N8.log 2>&1; echo "rc=$?" >> gpurun_out/r02z_bench_N8.log; tail -n 2 gpurun_out/r02z_bench_N8.log | cut -c1-900
