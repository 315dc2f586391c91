#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PRE=supersonic_b200/lib/segv_trace.so
LD_PRELOAD=$PRE timeout 900 python -m pytest tests/test_multi_gpu_nccl.py -m gpu -q -p no:cacheprovider -k "c_abi or hash_join" > gpurun_out/tests_nccl_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/tests_nccl_2gpu.log
tail -n 8 gpurun_out/tests_nccl_2gpu.log
SSB200_DEBUG_SHARD_JOIN=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/micro_shard_join.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -24
