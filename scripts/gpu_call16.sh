#!/bin/bash
# round-2 GPU call 16: lin_iter after the load / reduce / product trimming, mid-d batch lanes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lin_iter_gpu.py tests/test_small_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -rf -s \
  -k "lin_iter or logdet_inv" > gpurun_out/c16_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c16_pytest.log
timeout 600 python -m pytest tests/test_large_gpu.py tests/test_scale_gpu.py -q -m gpu --no-header -p no:cacheprovider -rf \
  -k "minimize_stages or backtracking or c2_logistic or beyond_onchip" > gpurun_out/c16_pytest_large.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c16_pytest_large.log
DAGMA_B200_LIB=build/variants/libdagma_ltrace.so timeout 200 python scripts/lin_trace.py logistic 100 10000 > gpurun_out/c16_ltrace.log 2>&1
DAGMA_B200_LIB=build/variants/libdagma_ltrace.so timeout 200 python scripts/lin_trace.py l2 100 400 >> gpurun_out/c16_ltrace.log 2>&1
DAGMA_B200_LIB=build/variants/libdagma_ltrace.so timeout 200 python scripts/lin_trace.py l2 128 512 >> gpurun_out/c16_ltrace.log 2>&1
timeout 300 python scripts/perf_c2c3.py > gpurun_out/c16_perf.log 2>&1
timeout 600 python scripts/perf_midd_batch.py 100 16 3000 > gpurun_out/c16_batch.log 2>&1
cat gpurun_out/c16_ltrace.log gpurun_out/c16_perf.log gpurun_out/c16_batch.log
grep -E "passed|failed|error|mid-d batch|rc=" gpurun_out/c16_pytest.log gpurun_out/c16_pytest_large.log | tail
