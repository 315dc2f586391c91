#!/bin/bash
# One GPU-box session: crash repro with native backtrace, the GPU suite in separate processes
# (so a fault in one part cannot void the others), then both bench arms.
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PRE=supersonic_b200/lib/segv_trace.so
LD_PRELOAD=$PRE timeout 300 python tools/repro_join_types.py > gpurun_out/repro_join.log 2>&1; echo "repro rc=$?" >> gpurun_out/repro_join.log
LD_PRELOAD=$PRE timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -k "not unsupported_join_types and not vectors_late" > gpurun_out/tests_main.log 2>&1; echo "rc=$?" >> gpurun_out/tests_main.log
LD_PRELOAD=$PRE timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "vectors_late" > gpurun_out/tests_late.log 2>&1; echo "rc=$?" >> gpurun_out/tests_late.log
LD_PRELOAD=$PRE timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -k "unsupported_join_types" > gpurun_out/tests_jointypes.log 2>&1; echo "rc=$?" >> gpurun_out/tests_jointypes.log
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
tail -5 gpurun_out/repro_join.log gpurun_out/tests_main.log gpurun_out/tests_late.log gpurun_out/tests_jointypes.log gpurun_out/bench.log
