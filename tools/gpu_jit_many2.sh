#!/bin/bash
export SSB200_JIT_MANY_GROUPS=9
for g in 4000 1000 300 64 16; do timeout 300 python tools/micro_jit_many.py 200000000 $g 2>&1 | tail -n 2; done
