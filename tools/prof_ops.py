"""One short run of one operator for ncu captures: python tools/prof_ops.py group|group_small|sort|join [rows]."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.getcwd())
sys.argv = [sys.argv[0]] + sys.argv[1:]
what = sys.argv[1] if len(sys.argv) > 1 else "group"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000_000
import importlib.util
spec = importlib.util.spec_from_file_location("micro_ops", os.path.join(os.path.dirname(os.path.abspath(__file__)), "micro_ops.py"))
m = importlib.util.module_from_spec(spec)
spec.loader.exec_module(m)
if what == "group":
    m.group_bench(rows, 1_000_000, "C3 group-by SUM(double)+COUNT")
elif what == "group_small":
    m.group_bench(rows, 6, "group-by 6 groups")
elif what == "sort":
    m.sort_bench(rows)
elif what == "join":
    m.join_bench(rows // 10, rows)
elif what == "partition":
    m.partition_bench(rows, 8)
