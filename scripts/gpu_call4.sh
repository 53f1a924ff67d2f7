#!/bin/bash
# 2-GPU call: multi-GPU parity check + the bench line at N = 2 (sharded C2 / C3, strong scaling)
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/mgpu_check.py > gpurun_out/c4_mgpu_check.log 2>&1
echo "mgpu_check rc=$?" >> gpurun_out/c4_mgpu_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c4_bench_n2.json 2> gpurun_out/c4_bench_n2.err
echo "bench rc=$?" >> gpurun_out/c4_bench_n2.err
tail -3 gpurun_out/c4_mgpu_check.log; tail -3 gpurun_out/c4_bench_n2.err; tail -c 1500 gpurun_out/c4_bench_n2.json
