#!/bin/bash
# round-2 GPU call 36: ncu evidence of the CURRENT blocked inverse (after the tile-step prefetch) and of the bench kernel
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:outer_step -s 20 -c 1 -f -o gpurun_out/prof_outer_step_r2b \
  python scripts/prof_inverse.py 2000 > gpurun_out/c36_ncu_outer.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c5_iter_r2b.csv \
  python scripts/prof_c5_iter.py 2000 > gpurun_out/c36_ncu_c5iter.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r2b.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-full-fit --no-c5 > gpurun_out/c36_ncu_launches.log 2>&1
tail -2 gpurun_out/c36_ncu_outer.log; tail -2 gpurun_out/c36_ncu_c5iter.log; tail -2 gpurun_out/c36_ncu_launches.log
