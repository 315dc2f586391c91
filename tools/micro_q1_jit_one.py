"""One Q1-shape pass through the run-time compiled kernel (for ncu). usage: python tools/micro_q1_jit_one.py [rows]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SSB200_GROUP_JIT", "1")
import bench
from supersonic_b200 import capi
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 67_000_000
ctx = capi.Context(0)
r = bench.q1_aux(capi, ctx, 0, 1, rows, None, None)
print("rows=%d  %.3f ms  %.2f G rows/s" % (rows, r["seconds"] * 1e3, r["value"] / 1e9))
