#!/bin/bash
# Q1 shape: the aggregation sink under different tile variants / resident CTAs.
mkdir -p gpurun_out
: > gpurun_out/q1_variants.log
for cfg in "6 2" "6 3" "5 2" "5 3" "0 1" "1 1" "7 2" "8 2" "8 3"; do
  set -- $cfg
  echo "== variant $1 ctas $2" >> gpurun_out/q1_variants.log
  SSB200_SINK_VARIANT=$1 SSB200_SINK_CTAS=$2 SSB200_DEBUG_PLAN=1 timeout 300 python tools/micro_q1_sink.py 200000000 2>&1 | grep -v "column groups" | tail -n 3 >> gpurun_out/q1_variants.log
done
cat gpurun_out/q1_variants.log
