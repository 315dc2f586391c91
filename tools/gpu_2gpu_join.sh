#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PRE=supersonic_b200/lib/segv_trace.so
LD_PRELOAD=$PRE timeout 900 python -m pytest tests/test_device_gpu.py tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -k "materialize or apply_to_children or join" > gpurun_out/tests_join_1gpu.log 2>&1; echo "rc=$?" >> gpurun_out/tests_join_1gpu.log
LD_PRELOAD=$PRE timeout 900 python -m pytest tests/test_multi_gpu_nccl.py -m gpu -q -p no:cacheprovider -k "c_abi" > gpurun_out/tests_nccl_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/tests_nccl_2gpu.log
SSB200_DEBUG_SHARD_JOIN=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/micro_shard_join.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -12
timeout 300 python tools/micro_ops.py 2>&1 | tail -15
tail -n 15 gpurun_out/tests_join_1gpu.log; tail -n 15 gpurun_out/tests_nccl_2gpu.log
