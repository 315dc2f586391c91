"""Q1 shape (BASELINE config 5) on one GPU: the run-time compiled aggregation kernel (csrc/jit.cu) against the
interpreting sink, over launch shapes. usage: python tools/micro_q1_jit.py [rows]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from supersonic_b200 import capi
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
ctx = capi.Context(0)
os.environ["SSB200_DEBUG_PLAN"] = "1"
def run(tag):
    r = bench.q1_aux(capi, ctx, 0, 1, rows, None, None)
    print("%s rows=%d  %.3f ms  %.2f G rows/s  %.0f GB/s algorithmic" % (tag, rows, r["seconds"] * 1e3, r["value"] / 1e9, r["algorithmic_gbs_per_gpu"]), flush=True)
os.environ["SSB200_GROUP_JIT"] = "0"
run("interpreting sink")
os.environ["SSB200_GROUP_JIT"] = "1"
for t, r_, mc, pf in ((192, 1, "", 1), (256, 1, "", 1), (224, 1, "", 1), (192, 1, "", 2)):
    os.environ["SSB200_JIT_THREADS"], os.environ["SSB200_JIT_ROWS"], os.environ["SSB200_JIT_PREFETCH"] = str(t), str(r_), str(pf)
    if mc:
        os.environ["SSB200_JIT_MIN_CTAS"] = mc
    else:
        os.environ.pop("SSB200_JIT_MIN_CTAS", None)
    run("jit T=%d R=%d min_ctas=%s prefetch=%d" % (t, r_, mc or "auto", pf))
