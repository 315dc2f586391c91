#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PRE=supersonic_b200/lib/segv_trace.so
SSB200_DEBUG_PLAN=1 LD_PRELOAD=$PRE timeout 900 python -m pytest tests/test_device_gpu.py -m gpu -q -p no:cacheprovider -k "group" > gpurun_out/tests_group.log 2>&1; echo "rc=$?" >> gpurun_out/tests_group.log
timeout 300 python tools/micro_ops.py 2>&1 | head -4
SSB200_GROUP_DENSE=0 timeout 300 python tools/micro_ops.py 2>&1 | head -2
timeout 600 python tools/micro_group.py 2>&1 | tail -12
grep -v "column groups\|sink plan" gpurun_out/tests_group.log | tail -n 25
