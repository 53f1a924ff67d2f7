#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lin_iter_gpu.py tests/test_large_gpu.py -q -m gpu --no-header -p no:cacheprovider -rf \
  -k "side_by_side or minimize_batch or beyond_onchip" > gpurun_out/c39_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c39_pytest.log
timeout 300 python scripts/perf_midd_after_work.py > gpurun_out/c39_midd.log 2>&1
timeout 900 python bench.py > gpurun_out/c39_bench.json 2> gpurun_out/c39_bench.err
echo "bench rc=$?" >> gpurun_out/c39_bench.err
tail -3 gpurun_out/c39_pytest.log; cat gpurun_out/c39_midd.log; tail -2 gpurun_out/c39_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/c39_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['e2e']['value'])
print('c5', d['c5']['ms_per_iter'], d['c5']['inverse_ms'], d['c5']['e2e']['value'])
print('c2', d['c2']['us_per_iter'], d['c2']['e2e']['value'], 'c3', d['c3']['us_per_iter'], d['c3']['e2e']['value'])
print(d['mid_d_batch']['iters_per_s'], d['mid_d_batch']['lanes'])
P
