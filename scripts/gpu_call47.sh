#!/bin/bash
# round-2 GPU call 47 (4 GPUs): the bench line at N = 4 exactly as the driver launches it
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29588 \
  bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/c47_bench_n4.json 2> gpurun_out/c47_bench_n4.err
echo "rc=$?" >> gpurun_out/c47_bench_n4.err
tail -2 gpurun_out/c47_bench_n4.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/c47_bench_n4.json').read().strip().splitlines() if l.startswith('{')][-1])
print(sorted(d.keys()))
print(d['value'], d['e2e']['value'], d.get('strong_scaling'))
for r in d.get('c2_sharded',{}).get('by_n',[]): print({k:r[k] for k in ('n','us_per_iter_sharded','us_per_iter_one_gpu','speedup','sharded_path')})
for r in d.get('c3_sharded',{}).get('by_n',[]): print({k:r[k] for k in ('n','us_per_iter_sharded','us_per_iter_one_gpu','speedup','sharded_path')})
print(d.get('extras_error'), d.get('extras_timeout'))
P
