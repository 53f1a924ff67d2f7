#!/bin/bash
# round-2 GPU call 15: the persistent one-kernel iteration of DagmaLinear (d <= 128): parity, C2 timing fused / sequence,
# fit-kernel variant (filler skipped by the sub-partition mate of the diagonal warp), chain stamps of the blocked inverse
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lin_iter_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -rf -s > gpurun_out/c15_pytest_li.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c15_pytest_li.log
timeout 900 python -m pytest tests/test_large_gpu.py tests/test_scale_gpu.py -q -m gpu --no-header -p no:cacheprovider -rf \
  -k "minimize_stages or backtracking or c2_logistic or beyond_onchip" > gpurun_out/c15_pytest_large.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c15_pytest_large.log
for f in 1 0; do
  echo "== DAGMA_LIN_FUSED=$f" >> gpurun_out/c15_perf.log
  DAGMA_LIN_FUSED=$f timeout 300 python scripts/perf_c2c3.py >> gpurun_out/c15_perf.log 2>&1
done
timeout 600 python scripts/perf_fit_variants.py > gpurun_out/c15_variants.log 2>&1
DAGMA_B200_LIB=build/variants/libdagma_otrace.so timeout 300 python scripts/outer_trace.py 2000 > gpurun_out/c15_otrace.log 2>&1
DAGMA_B200_LIB=build/variants/libdagma_ltrace.so timeout 200 python scripts/lin_trace.py logistic 100 10000 > gpurun_out/c15_ltrace.log 2>&1
DAGMA_B200_LIB=build/variants/libdagma_ltrace.so timeout 200 python scripts/lin_trace.py l2 100 400 >> gpurun_out/c15_ltrace.log 2>&1
DAGMA_B200_LIB=build/variants/libdagma_mtrace.so timeout 120 python scripts/mlp_trace.py > gpurun_out/c15_mtrace.log 2>&1
cat gpurun_out/c15_ltrace.log gpurun_out/c15_mtrace.log
tail -12 gpurun_out/c15_pytest_li.log; tail -5 gpurun_out/c15_pytest_large.log; cat gpurun_out/c15_perf.log gpurun_out/c15_variants.log; head -16 gpurun_out/c15_otrace.log
