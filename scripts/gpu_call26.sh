#!/bin/bash
# round-2 GPU call 26 (8 GPUs): the peer-memory exchange of the row-sharded persistent kernel on more than two GPUs
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 \
  scripts/mgpu_peer_check.py > gpurun_out/c26_peer8.log 2>&1
echo "rc=$?" >> gpurun_out/c26_peer8.log
grep -E "parity|d=100|rc=|Error|error" gpurun_out/c26_peer8.log | head -20
