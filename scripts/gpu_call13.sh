#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mlp_gpu.py -q -m gpu --no-header -p no:cacheprovider -rf > gpurun_out/c13_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c13_pytest.log
DAGMA_B200_LIB=build/variants/libdagma_mtrace.so timeout 120 python scripts/mlp_trace.py > gpurun_out/c13_mlp_trace.log 2>&1
timeout 300 python scripts/perf_c2c3.py > gpurun_out/c13_perf.log 2>&1
tail -4 gpurun_out/c13_pytest.log; cat gpurun_out/c13_mlp_trace.log gpurun_out/c13_perf.log
