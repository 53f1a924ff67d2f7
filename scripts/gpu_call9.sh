#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/c9_perf.log
for h in 1 0; do
  echo "== DAGMA_CS_HEAD=$h" >> gpurun_out/c9_perf.log
  DAGMA_CS_HEAD=$h timeout 300 python scripts/perf_c5.py 2000 >> gpurun_out/c9_perf.log 2>&1
done
DAGMA_B200_LIB=build/variants/libdagma_otrace.so timeout 300 python scripts/outer_trace.py 2000 > gpurun_out/c9_otrace.log 2>&1
cat gpurun_out/c9_perf.log; head -14 gpurun_out/c9_otrace.log
