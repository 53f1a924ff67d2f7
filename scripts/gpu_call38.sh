#!/bin/bash
# round-2 GPU call 38: worker GEMMs of the persistent linear kernel balanced over the sub-partitions
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lin_iter_gpu.py tests/test_scale_gpu.py -q -m gpu --no-header -p no:cacheprovider -rf \
  -k "lin_iter or c2_logistic" > gpurun_out/c38_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c38_pytest.log
DAGMA_B200_LIB=build/variants/libdagma_ltrace.so timeout 200 python scripts/lin_trace.py logistic 100 10000 > gpurun_out/c38_ltrace.log 2>&1
timeout 300 python scripts/perf_c2c3.py > gpurun_out/c38_perf.log 2>&1
tail -4 gpurun_out/c38_pytest.log; cat gpurun_out/c38_ltrace.log gpurun_out/c38_perf.log
