#!/bin/bash
# round-2 GPU call 7: tile step with cp.async prefetch -- parity of the inverse paths, A/B timing, chain stamps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_large_gpu.py tests/test_scale_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -k "inv or large or c5 or variants or flow" > gpurun_out/c7_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c7_pytest.log
for lib in "" build/variants/libdagma_ts0.so; do
  echo "== lib=$lib" >> gpurun_out/c7_perf.log
  DAGMA_B200_LIB=$lib timeout 300 python scripts/perf_c5.py 2000 >> gpurun_out/c7_perf.log 2>&1
done
for v in otrace otrace0; do
  DAGMA_B200_LIB=build/variants/libdagma_$v.so timeout 300 python scripts/outer_trace.py 2000 > gpurun_out/c7_$v.log 2>&1
done
tail -3 gpurun_out/c7_pytest.log; cat gpurun_out/c7_perf.log; cat gpurun_out/c7_otrace.log
