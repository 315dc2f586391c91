#!/bin/bash
mkdir -p gpurun_out
PRE=supersonic_b200/lib/segv_trace.so
LD_PRELOAD=$PRE timeout 900 python -m pytest tests/test_device_gpu.py -m gpu -q -p no:cacheprovider -x -k "fused" > gpurun_out/tests_fused.log 2>&1; echo "rc=$?" >> gpurun_out/tests_fused.log
SSB200_DEBUG_PLAN=1 timeout 600 python tools/micro_q1.py 200000000 > gpurun_out/micro_q1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:expr_sink -s 1 -c 1 -o gpurun_out/prof_q1_sink -f python tools/micro_q1_sink.py 67000000 > gpurun_out/q1_prof.log 2>&1
tail -n 5 gpurun_out/tests_fused.log; grep -v "column groups" gpurun_out/micro_q1.log | tail -n 5
