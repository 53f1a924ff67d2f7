#!/bin/bash
# round-2 GPU call 17: lin_iter with the outputs leaving from the accumulators; batch lane sweep; ncu of the kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lin_iter_gpu.py tests/test_small_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -rf -s \
  -k "lin_iter or logdet_inv" > gpurun_out/c17_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c17_pytest.log
DAGMA_B200_LIB=build/variants/libdagma_ltrace.so timeout 200 python scripts/lin_trace.py logistic 100 10000 > gpurun_out/c17_ltrace.log 2>&1
DAGMA_B200_LIB=build/variants/libdagma_ltrace.so timeout 200 python scripts/lin_trace.py l2 100 400 >> gpurun_out/c17_ltrace.log 2>&1
timeout 300 python scripts/perf_c2c3.py > gpurun_out/c17_perf.log 2>&1
for l in 2 4 6; do DAGMA_BATCH_LANES=$l timeout 300 python scripts/perf_midd_batch.py 100 16 3000 2>&1 | head -1 >> gpurun_out/c17_batch.log; done
timeout 300 python scripts/perf_midd_batch.py 100 16 3000 >> gpurun_out/c17_batch.log 2>&1
timeout 300 python scripts/perf_midd_batch.py 128 16 3000 >> gpurun_out/c17_batch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_iter -c 1 -f -o gpurun_out/prof_lin_iter_c2_r2 \
  python scripts/prof_c2.py logistic 100 10000 300 > gpurun_out/c17_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c2_minimize_r2.csv \
  python scripts/prof_c2.py logistic 100 10000 1000 > gpurun_out/c17_ncu_launches.log 2>&1
cat gpurun_out/c17_ltrace.log gpurun_out/c17_perf.log gpurun_out/c17_batch.log
grep -E "passed|failed|error|rc=" gpurun_out/c17_pytest.log | tail -4; tail -3 gpurun_out/c17_ncu.log
