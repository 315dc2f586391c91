#!/bin/bash
# e2e throughput with cached vs streaming stores in the narrowing loop
for nt in 0 1 0 1; do
  SSB200_NARROW_NT=$nt timeout 300 python bench.py --rows 67108864 --group-rows 1000000 --join-probe-rows 1000000 \
    --q1-rows 1000000 --cpu-rows 1000000 --e2e-steps 4 2>/dev/null > /tmp/e2e_nt.json
  python - "$nt" <<'PY'
import json, sys
d = json.load(open("/tmp/e2e_nt.json"))
print("streaming stores", sys.argv[1], "e2e rows/s", d["e2e"]["value"])
PY
done
