#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mlp_gpu.py -q -m gpu --no-header -p no:cacheprovider -rf > gpurun_out/c11_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c11_pytest.log
rm -f gpurun_out/c11_perf.log
DAGMA_MLP_FUSED=1 timeout 300 python scripts/perf_c2c3.py >> gpurun_out/c11_perf.log 2>&1
tail -5 gpurun_out/c11_pytest.log; cat gpurun_out/c11_perf.log
