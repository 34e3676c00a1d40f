#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2gpu.log
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_1gpu.log
for f in gpus dist_check bench_2gpu bench_1gpu; do echo "== $f"; tail -n 6 gpurun_out/$f.*; done
