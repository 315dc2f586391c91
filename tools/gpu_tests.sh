#!/bin/bash
# The GPU suite in separate processes (a fault in one part cannot void the others), native backtrace on a crash.
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PRE=supersonic_b200/lib/segv_trace.so
LD_PRELOAD=$PRE timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/tests_parity.log 2>&1; echo "rc=$?" >> gpurun_out/tests_parity.log
LD_PRELOAD=$PRE timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --ignore=tests/test_parity_gpu.py > gpurun_out/tests_rest.log 2>&1; echo "rc=$?" >> gpurun_out/tests_rest.log
tail -n 15 gpurun_out/tests_parity.log; tail -n 6 gpurun_out/tests_rest.log
