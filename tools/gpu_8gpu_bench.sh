#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_8gpu.log 2> gpurun_out/bench_8gpu.err; echo "rc=$?" >> gpurun_out/bench_8gpu.err
tail -n 5 gpurun_out/bench_8gpu.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_8gpu.log").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], json.dumps(d.get("scale_aux")))
    print(d["aux"]["hash_join"]["exchange"])
except Exception as e:
    print("no bench line", e)
PY
