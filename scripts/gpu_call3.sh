#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/perf_fit_variants.py > gpurun_out/c3_variants.log 2>&1
for lib in trace trace0; do for n in 296 148; do
  DAGMA_B200_LIB=build/variants/libdagma_$lib.so timeout 120 python scripts/sweep_trace.py $n > gpurun_out/c3_${lib}_$n.log 2>&1
done; done
timeout 1500 python -m pytest tests -q -m gpu -rf --no-header -p no:cacheprovider -x > gpurun_out/c3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
cat gpurun_out/c3_variants.log; tail -8 gpurun_out/c3_pytest.log
