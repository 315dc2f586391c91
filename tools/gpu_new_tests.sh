#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PRE=supersonic_b200/lib/segv_trace.so
LD_PRELOAD=$PRE timeout 900 python -m pytest tests/test_device_gpu.py -m gpu -q -p no:cacheprovider -k "attached or join" > gpurun_out/tests_new.log 2>&1; echo "rc=$?" >> gpurun_out/tests_new.log
tail -n 30 gpurun_out/tests_new.log
