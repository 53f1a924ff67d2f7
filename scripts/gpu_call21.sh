#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lin_iter_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -rf -s > gpurun_out/c21_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c21_pytest.log
timeout 900 python scripts/perf_c2_n.py > gpurun_out/c21_perf_n.log 2>&1
grep -E "passed|failed|rc=|Error" gpurun_out/c21_pytest.log | tail -5; grep "streamed\|20000\|15001\|30011\|12000" gpurun_out/c21_pytest.log | head; cat gpurun_out/c21_perf_n.log
