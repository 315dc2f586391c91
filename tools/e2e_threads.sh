#!/bin/bash
# e2e throughput of the streamed filter+project plan against the number of host threads that narrow the chunks
nproc
for t in 8 16 32 64 128; do
  SSB200_HOST_THREADS=$t timeout 300 python bench.py --rows 67108864 --group-rows 1000000 --join-probe-rows 1000000 \
    --q1-rows 1000000 --cpu-rows 1000000 --e2e-steps 4 2>/dev/null > /tmp/e2e_$t.json
  python - "$t" <<'PY'
import json, sys
t = sys.argv[1]
d = json.load(open("/tmp/e2e_%s.json" % t))
print("threads", t, "e2e rows/s", d["e2e"]["value"], "h2d bytes/step", d["e2e"]["h2d_bytes_per_step"])
PY
done
