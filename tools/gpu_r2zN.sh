#!/bin/bash
# final code at N = $1 GPUs of one box: bench.py under torchrun (weak, strong, stream v1, e2e legs)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02z_bench_N$N.log 2>&1; echo "rc=$?" >> gpurun_out/r02z_bench_N$N.log; tail -n 2 gpurun_out/r02z_bench_N$N.log | cut -c1-300
