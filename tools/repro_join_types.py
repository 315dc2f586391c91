"""GPU-box repro of the round-1 crash: RIGHT_OUTER / FULL_OUTER hash join through the mirror.
Run under LD_PRELOAD=supersonic_b200/lib/segv_trace.so for a native backtrace."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from supersonic_b200 import ssplan as sp
lib = sp.PlanLib(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "supersonic_b200", "lib", "libssb200_plan.so"))
t = lambda n: [sp.Column("k", sp.INT64, np.arange(n)), sp.Column("v", sp.INT64, np.arange(n))]
for jt in ("INNER", "RIGHT_OUTER", "FULL_OUTER"):
    plan = "(hash_join %s (named k) (named k) (multi (0 (named v)) (1 (rename (v w)))) UNIQUE (scan 0) (scan 1))" % jt
    for nl, nr in ((3, 3), (3, 0)):
        print("running", jt, nl, nr, flush=True)
        r = lib.run(plan, [t(nl), t(nr)])
        print("  ->", r.code, r.error, r.rows, flush=True)
print("done")
