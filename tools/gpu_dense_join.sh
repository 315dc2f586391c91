#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PRE=supersonic_b200/lib/segv_trace.so
LD_PRELOAD=$PRE timeout 900 python -m pytest tests/test_device_gpu.py tests/test_parity_gpu.py tests/test_golden_fixtures.py -m gpu -q -p no:cacheprovider -k "join or file_scan" > gpurun_out/tests_join.log 2>&1; echo "rc=$?" >> gpurun_out/tests_join.log
tail -n 25 gpurun_out/tests_join.log
timeout 300 python tools/micro_ops.py 2>&1 | grep "^join"
SSB200_JOIN_DENSE=0 timeout 300 python tools/micro_ops.py 2>&1 | grep "^join" | head -1
timeout 300 python - <<'PY'
import sys; sys.path.insert(0, ".")
import bench
from supersonic_b200 import capi
ctx = capi.Context(0)
r = bench.hash_join_aux(capi, ctx, 0, 1, 125_000_000, 12_500_000, None, None)
print("C4 single GPU: %.3f ms" % (r["seconds"] * 1e3))
PY
