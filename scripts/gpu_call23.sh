#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/c23_bench.json 2> gpurun_out/c23_bench.err
echo "bench rc=$?" >> gpurun_out/c23_bench.err
timeout 200 python -m pytest tests/test_utils_gpu.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/c23_pytest.log 2>&1
tail -3 gpurun_out/c23_bench.err; tail -2 gpurun_out/c23_pytest.log; python - <<'P'
import json
d=json.loads(open('gpurun_out/c23_bench.json').read().strip().splitlines()[-1])
print(sorted(d.keys()))
print(d.get('mid_d_batch'))
print(d['value'], d['roofline']['frac'], d['c2']['us_per_iter'], d['c3']['us_per_iter'], d['c5']['inverse_ms'])
P
