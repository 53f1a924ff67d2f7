#!/bin/bash
# round-2 GPU call 20: programmatic dependent launch between the outer steps of the blocked inverse
mkdir -p gpurun_out
for p in 0 1; do
  echo "== DAGMA_PDL=$p" >> gpurun_out/c20_perf.log
  DAGMA_PDL=$p timeout 300 python scripts/perf_c5.py 2000 >> gpurun_out/c20_perf.log 2>&1
done
DAGMA_PDL=1 timeout 600 python -m pytest tests/test_scale_gpu.py tests/test_large_gpu.py -q -m gpu --no-header -p no:cacheprovider -rf \
  -k "c5_inverse or blocked_logdet or rider or c5_reduced" > gpurun_out/c20_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c20_pytest.log
cat gpurun_out/c20_perf.log; tail -4 gpurun_out/c20_pytest.log
