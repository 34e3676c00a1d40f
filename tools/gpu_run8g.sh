#!/bin/bash
# 8-GPU evidence: bit-identical sharded sums (tools/dist_check.py) and the weak-scaling bench at N = 8, 4, 2, 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/dist_check.py > gpurun_out/dist_check8.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check8.log
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_${n}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_${n}gpu.log
done
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu.log 2>&1
grep -h "rank" gpurun_out/dist_check8.log | cut -c1-160; for n in 8 4 2 1; do tail -n 2 gpurun_out/bench_${n}gpu.log | cut -c1-300; done
