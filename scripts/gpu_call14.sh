#!/bin/bash
# round-2 GPU call 14 (session 3): full GPU tests on the rebuilt library, bench line, perf of C5 pieces
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -rf --no-header -p no:cacheprovider --durations=15 > gpurun_out/c14_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c14_pytest.log
timeout 700 python bench.py > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err
echo "bench rc=$?" >> gpurun_out/c14_bench.err
timeout 300 python scripts/perf_c5.py 2000 > gpurun_out/c14_perf_c5.log 2>&1
tail -25 gpurun_out/c14_pytest.log; tail -3 gpurun_out/c14_bench.err; cat gpurun_out/c14_perf_c5.log
