#!/bin/bash
# round-2 GPU call 22: full GPU suite + bench line after the persistent linear kernel, peer exchange, graph metrics
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu -rf --no-header -p no:cacheprovider --durations=8 > gpurun_out/c22_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c22_pytest.log
timeout 800 python bench.py > gpurun_out/c22_bench.json 2> gpurun_out/c22_bench.err
echo "bench rc=$?" >> gpurun_out/c22_bench.err
tail -16 gpurun_out/c22_pytest.log; tail -3 gpurun_out/c22_bench.err; cut -c1-600 gpurun_out/c22_bench.json
