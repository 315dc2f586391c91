#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PRE=supersonic_b200/lib/segv_trace.so
nvidia-smi -L > gpurun_out/gpus.txt
LD_PRELOAD=$PRE timeout 1500 python -m pytest tests/test_multi_gpu_nccl.py -m gpu -q -p no:cacheprovider ${NCCL_TEST_FILTER:+-k "$NCCL_TEST_FILTER"} > gpurun_out/tests_nccl_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/tests_nccl_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; echo "rc=$?" >> gpurun_out/bench_2gpu.err
tail -n 30 gpurun_out/tests_nccl_2gpu.log; tail -n 5 gpurun_out/bench_2gpu.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_2gpu.log").read().strip().splitlines()[-1])
    print(json.dumps(d.get("scale_aux")), d["aux"]["hash_join"]["exchange"])
except Exception as e:
    print("no bench line", e)
PY
