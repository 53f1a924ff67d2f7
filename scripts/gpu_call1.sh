#!/bin/bash
# round-2 GPU call 1: full GPU test suite, phase trace of the DMMA fit kernel, default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -rf --no-header -p no:cacheprovider > gpurun_out/c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
for n in 296 148; do
  DAGMA_B200_LIB=build/variants/libdagma_trace.so timeout 120 python scripts/sweep_trace.py $n > gpurun_out/c1_trace_$n.log 2>&1
done
timeout 600 python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
echo "bench rc=$?" >> gpurun_out/c1_bench.err
tail -5 gpurun_out/c1_pytest.log
