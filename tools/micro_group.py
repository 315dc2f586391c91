"""Group-by micro-benchmarks only (A/B of the kernel variants via SSB200_GROUP_FAST / SSB200_GROUP_TINY)."""
import os
import sys
import importlib.util

spec = importlib.util.spec_from_file_location("micro_ops", os.path.join(os.path.dirname(os.path.abspath(__file__)), "micro_ops.py"))
m = importlib.util.module_from_spec(spec)
spec.loader.exec_module(m)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
print("FAST=%s TINY=%s" % (os.environ.get("SSB200_GROUP_FAST", "default"), os.environ.get("SSB200_GROUP_TINY", "default")))
m.group_bench(rows, 1_000_000, "C3 group-by SUM(double)+COUNT")
m.group_bench(rows, 1000, "group-by 1000 groups")
m.group_bench(rows, 6, "group-by 6 groups")
