"""Summarises one kernel of an .ncu-rep: key metrics, stall reasons and the hottest source lines.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [n_lines]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units = r[0], r[1]
for vals in r[2:]:
    d = dict(zip(hdr, vals))
    print("==", d.get("Kernel Name", "")[:80], "grid", d.get("launch__grid_size"), "block", d.get("launch__block_size"))
    for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
              "sm__inst_executed.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
              "launch__occupancy_limit_registers", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
              "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
              "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]:
        if k in d:
            print("  %-62s %s %s" % (k, d[k], units[hdr.index(k)]))
    st = []
    for k, v in d.items():
        if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
            try:
                st.append((float(v.replace(",", "")), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1
    print("  stalls:", ", ".join("%s %.0f%%" % (k, 100 * v / tot) for v, k in sorted(st, reverse=True)[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur, agg = None, []
for x in rows:
    if len(x) >= 2 and x[0] == "File Path":
        cur = x[1].split("/")[-1]
        continue
    if len(x) < 8 or x[0] in ("Line No", "Function Name", ""):
        continue
    try:
        agg.append((int(x[7]), int(x[4]), cur, x[0], x[1][:100]))
    except ValueError:
        pass
ti, ts = sum(a[0] for a in agg) or 1, sum(a[1] for a in agg) or 1
print("  source lines by instructions executed (total %d, samples %d):" % (ti, ts))
for a in sorted(agg, reverse=True)[:nl]:
    print("   %5.1f%% inst %5.1f%% smp  %s:%s  %s" % (100.0 * a[0] / ti, 100.0 * a[1] / ts, a[2], a[3], a[4]))
