#!/bin/bash
# round-2 GPU call 19 (2 GPUs): row-sharded persistent iteration with the cross-GPU sum over NVLink peer memory
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  scripts/mgpu_peer_check.py > gpurun_out/c19_peer.log 2>&1
echo "rc=$?" >> gpurun_out/c19_peer.log
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 200 python scripts/perf_midd_batch.py 128 16 3000 8,6 > gpurun_out/c19_batch.log 2>&1
grep -v "^W\|Warning\|warn" gpurun_out/c19_peer.log | tail -30; cat gpurun_out/c19_batch.log
