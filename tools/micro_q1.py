"""Q1 shape (BASELINE config 5) on one GPU: ssb_group_update_program with and without the aggregation sink.
usage: python tools/micro_q1.py [rows]"""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from supersonic_b200 import capi
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
ctx = capi.Context(0)
for sink in ("1", "0"):
    os.environ["SSB200_GROUP_SINK"] = sink
    r = bench.q1_aux(capi, ctx, 0, 1, rows, None, None)
    print("sink=%s rows=%d  %.3f ms  %.2f G rows/s  %.0f GB/s algorithmic (two-step whole table %.3f ms)"
          % (sink, rows, r["seconds"] * 1e3, r["value"] / 1e9, r["algorithmic_gbs_per_gpu"], r["whole_table_two_step_seconds"] * 1e3), flush=True)
