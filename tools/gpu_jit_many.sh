#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PRE=supersonic_b200/lib/segv_trace.so
LD_PRELOAD=$PRE timeout 900 python -m pytest tests/test_device_gpu.py -m gpu -q -p no:cacheprovider -x -k "fused" > gpurun_out/tests_fused.log 2>&1; echo "rc=$?" >> gpurun_out/tests_fused.log
tail -n 4 gpurun_out/tests_fused.log
SSB200_DEBUG_PLAN=1 timeout 600 python tools/micro_jit_many.py 200000000 1000000 2>&1 | grep -v "sink plan\|column groups" | tail -n 8
timeout 600 python tools/micro_jit_many.py 200000000 20000 2>&1 | tail -n 3
