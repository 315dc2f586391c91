#!/bin/bash
# the adaptive switch of transfer narrowing: few host threads must not cost throughput
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "streaming or c1" 2>&1 | tail -1
for t in 4 default; do
  if [ "$t" = default ]; then unset SSB200_HOST_THREADS; else export SSB200_HOST_THREADS=$t; fi
  timeout 300 python bench.py --rows 67108864 --group-rows 1000000 --join-probe-rows 1000000 --q1-rows 1000000 \
    --cpu-rows 1000000 --e2e-steps 3 2>/dev/null > /tmp/e2e_ad.json
  python - "$t" <<'PY'
import json, sys
d = json.load(open("/tmp/e2e_ad.json"))
print("host threads", sys.argv[1], "e2e rows/s", d["e2e"]["value"], "h2d bytes/step", d["e2e"]["h2d_bytes_per_step"])
PY
done
