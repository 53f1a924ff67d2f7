#!/bin/bash
# round-2 GPU call 8: CS strip formed at the head of the outer step -- parity, A/B timing, chain stamps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_large_gpu.py tests/test_scale_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -k "inv or large or c5 or variants or flow" > gpurun_out/c8_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c8_pytest.log
rm -f gpurun_out/c8_perf.log
for h in 1 0; do
  echo "== DAGMA_CS_HEAD=$h" >> gpurun_out/c8_perf.log
  DAGMA_CS_HEAD=$h timeout 300 python scripts/perf_c5.py 2000 >> gpurun_out/c8_perf.log 2>&1
done
echo "== DAGMA_CS_HEAD=1 DAGMA_OUTER_SLEEP=0" >> gpurun_out/c8_perf.log
DAGMA_OUTER_SLEEP=0 timeout 300 python scripts/perf_c5.py 2000 >> gpurun_out/c8_perf.log 2>&1
DAGMA_B200_LIB=build/variants/libdagma_otrace.so timeout 300 python scripts/outer_trace.py 2000 > gpurun_out/c8_otrace.log 2>&1
tail -3 gpurun_out/c8_pytest.log; cat gpurun_out/c8_perf.log; head -24 gpurun_out/c8_otrace.log
