#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/dist_check.py > gpurun_out/dist_check2.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref_2gpu.log
grep -h "rank\|rc=" gpurun_out/dist_check2.log | cut -c1-200; tail -n 2 gpurun_out/bench_2gpu.log | cut -c1-600; tail -n 2 gpurun_out/bench_ref_2gpu.log | cut -c1-400
