#!/bin/bash
mkdir -p gpurun_out
DAGMA_B200_LIB=build/variants/libdagma_ltrace.so timeout 200 python scripts/lin_trace.py logistic 100 10000 > gpurun_out/c18_ltrace.log 2>&1
timeout 300 python scripts/perf_midd_batch.py 128 16 3000 1,2,4,8 > gpurun_out/c18_batch.log 2>&1
timeout 300 python scripts/perf_midd_batch.py 72 16 3000 1,2,4,8 >> gpurun_out/c18_batch.log 2>&1
cat gpurun_out/c18_ltrace.log gpurun_out/c18_batch.log
