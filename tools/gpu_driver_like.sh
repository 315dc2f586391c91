#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/tests_driver_like.log 2>&1; echo "rc=$?" >> gpurun_out/tests_driver_like.log
tail -n 6 gpurun_out/tests_driver_like.log
