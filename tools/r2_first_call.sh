#!/bin/bash
# First GPU call of the next round: everything that was written after the last GPU minute of round 1 was spent.
#   gpurun --timeout 1500 -- 'bash tools/r2_first_call.sh'
# Writes gpurun_out/r2_first_*.{log,txt,json}. Each step has its own timeout; a failing step does not stop the rest.
mkdir -p gpurun_out
O=gpurun_out/r2_first
# 1. the GPU suite (new since the last run: sort_signed_zero, group/join_signed_zero_keys golden cases)
(timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread 2>&1 | tail -15) > ${O}_tests.log 2>&1
# 2. ShardedSort through the real kernels on one GPU (world size 1 over NCCL)
(timeout 120 python tools/check_sharded_sort.py) > ${O}_sharded_sort.txt 2>&1
# 3. e2e with the AVX2 narrowing loop: default threads, then a thread sweep
(timeout 600 python bench.py) > ${O}_bench.json 2> ${O}_bench.err
(timeout 600 bash tools/e2e_threads.sh) > ${O}_e2e_threads.txt 2>&1
# 4. micro benchmarks (regression check against profiles/r1z_micro_ops.txt, r1s_micro_expr.txt)
(timeout 300 python tools/micro_expr.py 400000000) > ${O}_micro_expr.txt 2>&1
(timeout 300 python tools/micro_ops.py 200000000) > ${O}_micro_ops.txt 2>&1
tail -n 5 ${O}_tests.log; cat ${O}_sharded_sort.txt | tail -n 3; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_first_bench.json"))
    print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("bench:", e)
PY
cat ${O}_e2e_threads.txt | tail -n 8
