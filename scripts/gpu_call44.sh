#!/bin/bash
# round-2 GPU call 44: final full GPU suite + final bench line (host pipeline in e2e)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/c44_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c44_pytest.log
timeout 900 python bench.py > gpurun_out/c44_bench.json 2> gpurun_out/c44_bench.err
echo "bench rc=$?" >> gpurun_out/c44_bench.err
tail -3 gpurun_out/c44_pytest.log; tail -2 gpurun_out/c44_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/c44_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])
print('c5', d['c5']['ms_per_iter'], d['c5']['inverse_ms'], d['c5']['e2e']['value'])
print('c2', d['c2']['us_per_iter'], d['c2']['e2e']['value'], 'c3', d['c3']['us_per_iter'], d['c3']['e2e']['value'])
print(d['mid_d_batch']['iters_per_s'], d['parity_sample']['identical_edge_sets'], d['c5']['parity']['edge_set_distance'])
P
