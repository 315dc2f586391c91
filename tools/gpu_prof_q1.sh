#!/bin/bash
mkdir -p gpurun_out
SSB200_DEBUG_PLAN=1 timeout 600 python tools/micro_q1.py 200000000 2>&1 | grep -v "column groups" > gpurun_out/micro_q1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:expr_sink -s 1 -c 1 -o gpurun_out/prof_q1_sink_r2d -f python tools/micro_q1_sink.py 67000000 > gpurun_out/q1_prof.log 2>&1
tail -n 5 gpurun_out/micro_q1.log
