#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PRE=supersonic_b200/lib/segv_trace.so
LD_PRELOAD=$PRE timeout 900 python -m pytest tests/test_device_gpu.py -m gpu -q -p no:cacheprovider -x -k "fused" > gpurun_out/tests_fused.log 2>&1; echo "rc=$?" >> gpurun_out/tests_fused.log
tail -n 4 gpurun_out/tests_fused.log
LD_PRELOAD=$PRE timeout 900 python -m pytest tests/test_reference_clients.py tests/test_strings_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/tests_clients.log 2>&1; echo "rc=$?" >> gpurun_out/tests_clients.log
tail -n 30 gpurun_out/tests_clients.log
timeout 600 python tools/micro_q1_jit.py 200000000 > gpurun_out/micro_q1_jit.log 2>&1
grep -v "sink plan\|column groups\|jit: compiled" gpurun_out/micro_q1_jit.log | tail -n 40
