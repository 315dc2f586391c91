#!/bin/bash
# One 1-GPU session: smoke(), the GPU suite in separate processes, then both bench arms.
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PRE=supersonic_b200/lib/segv_trace.so
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
LD_PRELOAD=$PRE timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/tests_parity.log 2>&1; echo "rc=$?" >> gpurun_out/tests_parity.log
LD_PRELOAD=$PRE timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --ignore=tests/test_parity_gpu.py > gpurun_out/tests_rest.log 2>&1; echo "rc=$?" >> gpurun_out/tests_rest.log
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
tail -n 3 gpurun_out/smoke.log; tail -n 4 gpurun_out/tests_parity.log gpurun_out/tests_rest.log gpurun_out/bench.err
