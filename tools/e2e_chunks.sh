#!/bin/bash
# e2e throughput of the streamed filter+project plan against the chunk size of the host-table stream
nproc
for c in 2097152 4194304 8388608 16777216; do
  SSB200_CHUNK_ROWS=$c timeout 300 python bench.py --rows 268435456 --group-rows 1000000 --join-probe-rows 1000000 \
    --q1-rows 1000000 --sort-rows 1000000 --cpu-rows 1000000 --e2e-steps 4 --steps 3 2>/dev/null > /tmp/e2e_$c.json
  python - "$c" <<'PY'
import json, sys
c = sys.argv[1]
d = json.loads(open("/tmp/e2e_%s.json" % c).read().strip().splitlines()[-1])
print("chunk rows", c, "e2e rows/s %.4g" % d["e2e"]["value"], "h2d bytes/step", d["e2e"]["h2d_bytes_per_step"])
PY
done
