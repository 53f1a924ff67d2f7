#!/bin/bash
# round-2 GPU call 24 (2 GPUs): bench line at N = 2 (strong scaling, c2_sharded with the peer kernel, c3_sharded)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 > gpurun_out/c24_bench_n2.json 2> gpurun_out/c24_bench_n2.err
echo "rc=$?" >> gpurun_out/c24_bench_n2.err
timeout 300 python -m pytest tests/test_lin_iter_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "minimize_batch or side_by_side" > gpurun_out/c24_pytest.log 2>&1
tail -2 gpurun_out/c24_bench_n2.err; tail -3 gpurun_out/c24_pytest.log
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/c24_bench_n2.json').read().strip().splitlines() if l.startswith('{')][-1])
print(sorted(d.keys()))
print(d['value'], d['e2e']['value'], d.get('strong_scaling'))
for r in d.get('c2_sharded',{}).get('by_n',[]): print(r)
print(d.get('c3_sharded')); print(d.get('extras_error'), d.get('extras_timeout'))
P
