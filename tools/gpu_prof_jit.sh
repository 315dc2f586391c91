#!/bin/bash
mkdir -p gpurun_out
export SSB200_JIT_THREADS=192 SSB200_JIT_ROWS=2 SSB200_JIT_MIN_CTAS=3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ssb_jit_rows -s 3 -c 1 -o gpurun_out/prof_q1_jit_r2h -f python tools/micro_q1_jit_one.py 67000000 > gpurun_out/q1_jit_prof.log 2>&1
tail -n 5 gpurun_out/q1_jit_prof.log
