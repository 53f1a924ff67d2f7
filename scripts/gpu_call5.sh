#!/bin/bash
# round-2 GPU call 5: full tests (telemetry on chip, residency), C5 path snapshots, bench line, ncu evidence
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -rf --no-header -p no:cacheprovider > gpurun_out/c5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c5_pytest.log
timeout 600 python scripts/c5_path_snapshots.py 4000 > gpurun_out/c5_path_snapshots.txt 2> gpurun_out/c5_path_snapshots.err
timeout 600 python bench.py > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
echo "bench rc=$?" >> gpurun_out/c5_bench.err
# ncu evidence (each after its command ran clean above / in earlier calls)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r2.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-full-fit --no-c5 > gpurun_out/c5_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fit_small_dmma -c 1 -f -o gpurun_out/prof_fit_small_r2 \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-full-fit --no-c5 > gpurun_out/c5_ncu_fit.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:outer_step -s 20 -c 1 -f -o gpurun_out/prof_outer_step_r2 \
  python scripts/prof_inverse.py 2000 > gpurun_out/c5_ncu_outer.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c5_iter_r2.csv \
  python scripts/prof_c5_iter.py 2000 > gpurun_out/c5_ncu_c5iter.log 2>&1
tail -6 gpurun_out/c5_pytest.log; cat gpurun_out/c5_path_snapshots.txt | cut -c1-220
