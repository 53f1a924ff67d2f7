#!/bin/bash
# round-2 GPU call 10: one-kernel MLP iteration -- parity against the launch sequence and the reference fixtures, timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mlp_gpu.py -q -m gpu --no-header -p no:cacheprovider -rf > gpurun_out/c10_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c10_pytest.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_mlp_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "tiny or odd" > gpurun_out/c10_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/c10_memcheck.log
for f in 1 0; do
  echo "== DAGMA_MLP_FUSED=$f" >> gpurun_out/c10_perf.log
  DAGMA_MLP_FUSED=$f timeout 300 python scripts/perf_c2c3.py >> gpurun_out/c10_perf.log 2>&1
done
tail -15 gpurun_out/c10_pytest.log; tail -5 gpurun_out/c10_memcheck.log; cat gpurun_out/c10_perf.log
