#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/perf_fit_variants.py > gpurun_out/c6_variants.log 2>&1
for lib in trace trace_il0 trace_il2; do for n in 296 148; do
  DAGMA_B200_LIB=build/variants/libdagma_$lib.so timeout 120 python scripts/sweep_trace.py $n > gpurun_out/c6_${lib}_$n.log 2>&1
done; done
cat gpurun_out/c6_variants.log
