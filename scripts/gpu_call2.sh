#!/bin/bash
# round-2 GPU call 2: lean Adam pass + fixed tests
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -rf --no-header -p no:cacheprovider > gpurun_out/c2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
timeout 600 python scripts/debug_c2.py > gpurun_out/c2_debug_c2.log 2>&1
for n in 296 148; do
  DAGMA_B200_LIB=build/variants/libdagma_trace.so timeout 120 python scripts/sweep_trace.py $n > gpurun_out/c2_trace_$n.log 2>&1
done
timeout 600 python bench.py > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
echo "bench rc=$?" >> gpurun_out/c2_bench.err
tail -12 gpurun_out/c2_pytest.log
