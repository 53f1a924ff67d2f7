#!/bin/bash
# round-2 GPU call 27 (2 GPUs): peer exchange inside the persistent MLP kernel; MLP tests on one GPU
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp_gpu.py -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/c27_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c27_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29566 \
  scripts/mgpu_peer_check.py > gpurun_out/c27_peer.log 2>&1
echo "rc=$?" >> gpurun_out/c27_peer.log
tail -3 gpurun_out/c27_pytest.log; grep -E "parity|d=100|MLP|rc=|Error|error|assert" gpurun_out/c27_peer.log | head -30
