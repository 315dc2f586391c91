"""One C4-shape join (125M x 12.5M, dense keys) for ncu. usage: python tools/micro_join_c4_one.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from supersonic_b200 import capi
ctx = capi.Context(0)
r = bench.hash_join_aux(capi, ctx, 0, 1, 125_000_000, 12_500_000, None, None)
print("C4 single GPU: %.3f ms" % (r["seconds"] * 1e3))
