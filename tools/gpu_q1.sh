#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PRE=supersonic_b200/lib/segv_trace.so
LD_PRELOAD=$PRE timeout 900 python -m pytest tests/test_device_gpu.py -m gpu -q -p no:cacheprovider -x -k "fused" > gpurun_out/tests_fused.log 2>&1; echo "rc=$?" >> gpurun_out/tests_fused.log
LD_PRELOAD=$PRE timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -k "unsupported_join_types or q1_shape or group_over_filter" > gpurun_out/tests_misc.log 2>&1; echo "rc=$?" >> gpurun_out/tests_misc.log
timeout 600 python tools/micro_q1.py 200000000 > gpurun_out/micro_q1.log 2>&1
tail -4 gpurun_out/tests_fused.log gpurun_out/tests_misc.log gpurun_out/micro_q1.log
