#!/bin/bash
# STRING / BINARY slice on the GPU, then the rest of the suite for regressions.
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PRE=supersonic_b200/lib/segv_trace.so
LD_PRELOAD=$PRE timeout 900 python -m pytest tests/test_strings_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/tests_strings.log 2>&1; echo "rc=$?" >> gpurun_out/tests_strings.log

tail -n 40 gpurun_out/tests_strings.log
