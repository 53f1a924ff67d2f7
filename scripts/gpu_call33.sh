#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lin_iter_gpu.py -q -m gpu --no-header -p no:cacheprovider -rf > gpurun_out/c33_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c33_pytest.log
timeout 900 python bench.py > gpurun_out/c33_bench.json 2> gpurun_out/c33_bench.err
echo "bench rc=$?" >> gpurun_out/c33_bench.err
tail -5 gpurun_out/c33_pytest.log; tail -2 gpurun_out/c33_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/c33_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['e2e']['value'])
print('c5', d['c5']['ms_per_iter'], d['c5']['inverse_ms'], d['c5']['e2e']['value'])
print('c2', d['c2']['us_per_iter'], 'c3', d['c3']['us_per_iter'])
print(d['mid_d_batch']['iters_per_s'], d['mid_d_batch']['lanes'])
P
