#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PRE=supersonic_b200/lib/segv_trace.so
LD_PRELOAD=$PRE timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -k "hybrid or distinct" > gpurun_out/tests_hybrid.log 2>&1; echo "rc=$?" >> gpurun_out/tests_hybrid.log
tail -n 5 gpurun_out/tests_hybrid.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ssb_jit_rows -s 3 -c 1 -o gpurun_out/prof_q1_jit_r2k -f python tools/micro_q1_jit_one.py 67000000 > gpurun_out/q1_jit_prof.log 2>&1
tail -n 2 gpurun_out/q1_jit_prof.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:join_probe_unique -s 1 -c 1 -o gpurun_out/prof_join_dense_r2k -f python tools/micro_join_c4_one.py > gpurun_out/join_prof.log 2>&1
tail -n 2 gpurun_out/join_prof.log
