#!/bin/bash
# multi-GPU evidence at N = $1 GPUs: bit-identity (one process per GPU and one process for all), C4 / C5 scaling pieces, bench
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_N$N.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 tools/dist_check.py > gpurun_out/dist_check_N$N.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check_N$N.log
grep -h "rank\|rc=" gpurun_out/dist_check_N$N.log | cut -c1-170 | tail -n 3
timeout 600 python -m pytest tests/test_gpu_multidevice.py -m gpu -q -k "sharded" > gpurun_out/pytest_multidevice_N$N.log 2>&1; tail -n 2 gpurun_out/pytest_multidevice_N$N.log
timeout 900 $TR --master-port 29522 tools/multi_gpu_suite.py > gpurun_out/suite_ranks_N$N.log 2>&1; echo "rc=$?" >> gpurun_out/suite_ranks_N$N.log; tail -n 2 gpurun_out/suite_ranks_N$N.log | cut -c1-400
timeout 900 python tools/multi_gpu_suite.py --single $N > gpurun_out/suite_single_N$N.log 2>&1; echo "rc=$?" >> gpurun_out/suite_single_N$N.log; tail -n 2 gpurun_out/suite_single_N$N.log | cut -c1-400
timeout 900 $TR --master-port 29523 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_N$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_N$N.log; tail -n 2 gpurun_out/bench_N$N.log | cut -c1-700
