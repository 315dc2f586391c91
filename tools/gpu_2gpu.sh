#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PRE=supersonic_b200/lib/segv_trace.so
nvidia-smi -L > gpurun_out/gpus.txt
LD_PRELOAD=$PRE timeout 1500 python -m pytest tests/test_multi_gpu_nccl.py -m gpu -q -p no:cacheprovider -x > gpurun_out/tests_nccl_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/tests_nccl_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; echo "rc=$?" >> gpurun_out/bench_2gpu.err
tail -n 12 gpurun_out/tests_nccl_2gpu.log; tail -n 3 gpurun_out/bench_2gpu.err; head -c 1500 gpurun_out/bench_2gpu.log
