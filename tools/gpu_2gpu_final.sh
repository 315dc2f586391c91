#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PRE=supersonic_b200/lib/segv_trace.so
LD_PRELOAD=$PRE timeout 1500 python -m pytest tests/test_multi_gpu_nccl.py -m gpu -q -p no:cacheprovider > gpurun_out/tests_nccl_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/tests_nccl_2gpu.log
tail -n 6 gpurun_out/tests_nccl_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; echo "rc=$?" >> gpurun_out/bench_2gpu.err
tail -n 2 gpurun_out/bench_2gpu.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_2gpu.log").read().strip().splitlines()[-1])
    print(d["value"], json.dumps(d.get("scale_aux")))
except Exception as e:
    print("no bench line", e)
PY
