"""One pass of the Q1 shape through ssb_group_update_program with the sink (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from supersonic_b200 import capi
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 67_000_000
os.environ["SSB200_GROUP_SINK"] = "1"
ctx = capi.Context(0)
r = bench.q1_aux(capi, ctx, 0, 1, rows, None, None)
print(r["seconds"])
