#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv \
  python bench.py --steps 2 --warmup 1 --group-rows 1000000 --q1-rows 1000000 --join-probe-rows 1000000 --sort-rows 1000000 --cpu-rows 1000000 --e2e-rows 8388608 --e2e-steps 1 > gpurun_out/launch_bench.log 2>&1
tail -n 2 gpurun_out/launch_bench.log | cut -c1-300
grep -c "expr_kernel" gpurun_out/launches_r2.csv; grep "expr_kernel" gpurun_out/launches_r2.csv | head -5 | cut -c1-250
