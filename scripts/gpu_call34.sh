#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lin_iter_gpu.py tests/test_large_gpu.py tests/test_scale_gpu.py tests/test_telemetry_trek_gpu.py -q -m gpu --no-header \
  -p no:cacheprovider -rf -k "backtracking or minimize_stages or c2_logistic or c5_reduced or trajectory or telemetry or d200" > gpurun_out/c34_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c34_pytest.log
timeout 300 python scripts/prof_c3.py 300 > gpurun_out/c34_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_iter -c 1 -f -o gpurun_out/prof_mlp_iter_c3_r2 \
  python scripts/prof_c3.py 300 > gpurun_out/c34_ncu.log 2>&1
tail -4 gpurun_out/c34_pytest.log; cat gpurun_out/c34_c3.log; tail -2 gpurun_out/c34_ncu.log
