#!/usr/bin/env python
"""bench.py -- filter+project (BASELINE config 2) on N B200s, one rank per GPU.

Workload (SURVEY.md section 8d, config C2 variant A): synthetic 8 x INT64 table generated
in HBM by the counter-based generator (seed 42); the plan is
    Filter(d < 2^19, project e, Compute({e := a*b + c, d}, ScanView(table)))
One "step" = one pass of the fused kernel over the rank's whole shard (rows per GPU fixed:
weak scaling; Compute/Filter rows are independent, so ranks share nothing and no collective
is on the data path).

Printed JSON (rank 0): value = rows/s over all ranks with inputs resident in HBM;
roofline = algorithmic bytes (32 B read + 8 B written per kept row) / kernel time against the
measured HBM peak; e2e = the same plan through the supersonic.h mirror with pinned HOST
buffers (H2D + kernel + D2H inside the timed region); cpu_baseline = the unmodified
reference (oracle/_ref) on a bounded sample, single thread (the engine is single-threaded).
--impl reference times the reference's own CPU path on all host cores instead.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PLAN = ("(filter (less (col d) (i64 524288)) (named e) (compute (compound "
        "(as e (plus (multiply (col a) (col b)) (col c))) (col d)) (scan 0)))")
SEED = 42
K_SEL = 1 << 19
# Both arms print exactly these (the driver divides the two lines only when they agree).
METRIC = "rows/sec, filter+project (Compute e=a*b+c, Filter d<2^19 project e) over 8xINT64 rows"


def workload(rows):
    return ("C2 variant A: Filter(d<2^19, project e, Compute(e=a*b+c, d)) over a %d-row 8xINT64 table per GPU "
            "(4 columns read, selectivity 0.5)" % rows)
# column generators: (kind, lo, span) -- a,b in [-2^31,2^31), c in [-2^62,2^62), d in [0,2^20)
GEN = {"a": (0, -(1 << 31), 1 << 32), "b": (0, -(1 << 31), 1 << 32), "c": (0, -(1 << 62), 1 << 63),
       "d": (0, 0, 1 << 20), "e": (0, 0, 0), "f": (0, 0, 0), "g": (0, 0, 0), "h": (0, 0, 0)}
COLS = "abcdefgh"


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def measured_traffic(rows):
    """DRAM bytes per launch from the committed ncu --set full capture (profiles/r1_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum at the row count it was taken on), scaled to this
    run's row count; None when no capture is committed."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        return float(t["dram_bytes"]) * rows / float(t["rows"])
    except Exception:
        return None


class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed region: an in-process NVML poller
    (every 2 ms; nvidia-smi -lms needs ~100 ms to deliver its first line, longer than ten
    6.7 ms steps), nvidia-smi as the fallback when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.lines, self.proc = device, [], None
        self.samples, self.stop_flag, self.thread, self.nvml = [], False, None, None
        self.window = [None, None]

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber: go through the PCI bus id of the CUDA device
            handle = None
            try:
                import torch
                bus = torch.cuda.get_device_properties(self.device).pci_bus_id
                dom = torch.cuda.get_device_properties(self.device).pci_domain_id
                dev = torch.cuda.get_device_properties(self.device).pci_device_id
                handle = pynvml.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (dom, bus, dev)).encode())
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            self.nvml = (pynvml, handle)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def mark(self, which):
        """which = 0: the timed region starts now; 1: it ended."""
        self.window[which] = time.perf_counter()

    def _poll(self):
        pynvml, h = self.nvml
        while not self.stop_flag:
            try:
                mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                reasons = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.samples.append((time.perf_counter(), float(mhz), int(reasons)))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = self.window
        inside = lambda t: (t0 is None or t >= t0) and (t1 is None or t <= t1)   # noqa: E731
        if self.nvml is not None:
            pynvml, _ = self.nvml
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            bits = {"hw_slowdown": pynvml.nvmlClocksEventReasonHwSlowdown,
                    "hw_thermal_slowdown": pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": pynvml.nvmlClocksEventReasonSwThermalSlowdown,
                    "sw_power_cap": pynvml.nvmlClocksEventReasonSwPowerCap}
            sel = [x for x in self.samples if inside(x[0])] or self.samples[-3:]
            reasons = sorted(nm for nm, b in bits.items() if any(x[2] & b for x in sel))
            sm = [x[1] for x in sel]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                    "samples": len(sm), "reasons": reasons, "source": "nvml poll inside the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        sel = [ln for t, ln in self.lines if inside(t)] or [ln for _, ln in self.lines[-3:]]
        for line in sel:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 20"}


def build_program(capi, ctx):
    """Bound DAG of the plan: e = a*b + c ; predicate d < K. Inputs: a, b, c, d."""
    n = capi.node
    I64, BOOL = capi.INT64, capi.BOOL
    nodes = [n(capi.OP_INPUT, I64, [0]), n(capi.OP_INPUT, I64, [1]), n(capi.OP_INPUT, I64, [2]),
             n(capi.OP_INPUT, I64, [3]), n(capi.OP_MUL, I64, [0, 1]), n(capi.OP_ADD, I64, [4, 2]),
             n(capi.OP_CONST, I64, [], i64=K_SEL), n(capi.OP_LT, BOOL, [3, 6])]
    return capi.Program(ctx, nodes, [I64] * 4, [0] * 4, [5], predicate=7)


def _cols(capi, items):
    arr = (capi.Column * max(1, len(items)))()
    for i, (d, n_, t) in enumerate(items):
        arr[i].data, arr[i].nulls, arr[i].dtype = d, n_, t
    return arr


def _timed(ctx, world, dist, torch, fn, repeats=3):
    """Host clock around device syncs (barrier before, sync after), max over ranks, best of
    `repeats` after one untimed run."""
    times = []
    for it in range(repeats + 1):
        ctx.sync()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        ctx.sync()
        if world > 1:
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        if it > 0:
            times.append(dt)
    return min(times), out


def group_by_aux(capi, ctx, rank, world, rows, dist, torch, comm=None):
    """BASELINE config 3 shape, sharded by row range: GroupAggregate(k; SUM(v), COUNT(*)) with
    k = u mod 1e6 (INT64), v = (u >> 44) * 2^-10 (exactly summable DOUBLE). Every rank aggregates
    its shard (ssb_group_update), the dense partial tables are all-gathered over NCCL and merged
    (ssb_group_merge). Returns rows/s over all ranks (host clock around device syncs)."""
    from supersonic_b200.distributed import merge_group_partials
    lib = ctx.lib
    first = rank * rows
    k = ctx.malloc(rows * 8 + 256)
    v = ctx.malloc(rows * 8 + 256)
    ctx.generate(k, rows, first, SEED, 20, 1, 0, 1000000)
    ctx.generate(v, rows, first, SEED, 21, 2, 0, 0)
    specs = (capi.AggSpec * 2)()
    specs[0].fn, specs[0].input, specs[0].in_type, specs[0].out_type = capi.AGG_SUM, 0, capi.DOUBLE, capi.DOUBLE
    specs[1].fn, specs[1].input, specs[1].in_type, specs[1].out_type = capi.AGG_COUNT, -1, capi.INT64, capi.UINT64
    kt, kn = (C.c_int32 * 1)(capi.INT64), (C.c_int32 * 1)(0)
    state = {}

    def once():
        g = C.c_void_p()
        ctx.check(lib.ssb_group_create(ctx.h, 1, kt, kn, 2, specs, 1000000, C.byref(g)))
        ctx.check(lib.ssb_group_update(g, _cols(capi, [(k, None, capi.INT64)]), _cols(capi, [(v, None, capi.DOUBLE)]), rows))
        n, ko, ao = merge_group_partials(ctx, g, [capi.INT64], [capi.DOUBLE, capi.UINT64], comm=comm)
        cnt = np.zeros(n, dtype=np.uint64)
        ctx.d2h(cnt, ao[1].data)
        state["groups"], state["rows_counted"] = n, int(cnt.sum())
        lib.ssb_group_destroy(g)

    best, _ = _timed(ctx, world, dist, torch, once, repeats=2)
    ctx.free(k)
    ctx.free(v)
    assert state["rows_counted"] == world * rows, "COUNT(*) over all groups must equal the table's rows"
    return {"metric": "rows/sec, GroupAggregate(k; SUM(v DOUBLE), COUNT(*)), 1M INT64 keys (BASELINE config 3 shape)",
            "value": world * rows / best, "unit": "rows/s", "rows_per_gpu": rows, "groups": int(state["groups"]),
            "seconds": best, "exchange": "ssb_shard_group_merge: partial groups hash-partitioned by key, one all-to-all, the owner "
                                         "merges, all-gather of the merged ranges (NCCL inside libssb200.so)" if world > 1 else "none",
            "algorithmic_gbs_per_gpu": rows * 16 / best / 1e9, "check": "sum of COUNT(*) == rows"}


def q1_aux(capi, ctx, rank, world, rows, dist, torch, comm=None):
    """BASELINE config 5 shape (TPC-H Q1 on synthetic lineitem columns), row-range
    sharded: Filter(ship <= 2450) -> Compute(disc_price, charge) as ONE fused kernel launch, then
    GroupAggregate({rf, ls}; 5 x SUM, COUNT(*)); per-rank partial tables merged after an
    all-gather (6 groups). Algorithmic bytes: 7 x 8 = 56 B per row."""
    from supersonic_b200.distributed import merge_group_partials
    lib = ctx.lib
    first = rank * rows
    I64, F64, BOOL = capi.INT64, capi.DOUBLE, capi.BOOL
    # inputs: qty price disc tax (DOUBLE, dyadic: sums and products stay exact), rf ls ship (INT64)
    gens = [(5, 16, 800), (5, 1600, 160000), (5, 0, 2), (5, 0, 2), (1, 0, 3), (1, 0, 2), (1, 0, 2500)]
    types = [F64, F64, F64, F64, I64, I64, I64]
    d_in = []
    for j, (kind, lo, span) in enumerate(gens):
        ptr = ctx.malloc(rows * 8 + 256)
        ctx.generate(ptr, rows, first, SEED, 30 + j, kind, lo, span)
        d_in.append(ptr)
    n = capi.node
    nodes = [n(capi.OP_INPUT, t, [j]) for j, t in enumerate(types)]                    # 0..6
    nodes += [n(capi.OP_CONST, F64, [], f64=1.0),                                      # 7
              n(capi.OP_SUB, F64, [7, 2]),                                             # 8: 1 - disc
              n(capi.OP_MUL, F64, [1, 8]),                                             # 9: disc_price
              n(capi.OP_ADD, F64, [7, 3]),                                             # 10: 1 + tax
              n(capi.OP_MUL, F64, [9, 10]),                                            # 11: charge
              n(capi.OP_CONST, I64, [], i64=2450),                                     # 12
              n(capi.OP_LE, BOOL, [6, 12])]                                            # 13: ship <= D
    # outputs: the keys rf, ls, then the aggregate inputs qty, price, disc_price, charge, disc
    prog = capi.Program(ctx, nodes, types, [0] * 7, [4, 5, 0, 1, 9, 11, 2], predicate=13)
    out_types = [I64, I64, F64, F64, F64, F64, F64]
    d_out = [ctx.malloc(rows * 8 + 256) for _ in out_types]
    d_count = ctx.malloc(8)
    specs = (capi.AggSpec * 6)()
    for i in range(5):   # SUM qty, price, disc_price, charge, disc
        specs[i].fn, specs[i].input, specs[i].in_type, specs[i].out_type = capi.AGG_SUM, i, F64, F64
    specs[5].fn, specs[5].input, specs[5].in_type, specs[5].out_type = capi.AGG_COUNT, -1, I64, capi.UINT64
    kt, kn = (C.c_int32 * 2)(I64, I64), (C.c_int32 * 2)(0, 0)
    state = {}

    in_cols = [(p_, None, t) for p_, t in zip(d_in, types)]

    def make_group():
        g = C.c_void_p()
        ctx.check(lib.ssb_group_create(ctx.h, 2, kt, kn, 6, specs, 6, C.byref(g)))
        return g

    def finish(g, tag):
        ng, ko, ao = merge_group_partials(ctx, g, [I64, I64], [F64] * 5 + [capi.UINT64], comm=comm)
        cnt = np.zeros(ng, dtype=np.uint64)
        ctx.d2h(cnt, ao[5].data)
        sums = np.zeros(ng, dtype=np.float64)
        ctx.d2h(sums, ao[3].data)
        state["groups"], state[tag] = ng, (int(cnt.sum()), float(sums.sum()))
        lib.ssb_group_destroy(g)

    def once():
        # the product path: one call; the library runs Filter -> Compute and the aggregation slice by slice
        g = make_group()
        ctx.check(lib.ssb_group_update_program(g, prog.h, _cols(capi, in_cols), rows))
        finish(g, "fused")

    def once_unfused():
        # for comparison: materialise the whole filtered / computed table, then aggregate it
        kept = prog.run_sync(in_cols, rows, [(p_, None, t) for p_, t in zip(d_out, out_types)])
        g = make_group()
        vals = [(d_out[j], None, F64) for j in range(2, 7)]
        ctx.check(lib.ssb_group_update(g, _cols(capi, [(d_out[0], None, I64), (d_out[1], None, I64)]), _cols(capi, vals), kept))
        state["kept"] = kept
        finish(g, "unfused")

    def jit_stats():
        k, ms, n = C.c_int64(), C.c_double(), C.c_int64()
        lib.ssb_jit_stats(C.byref(k), C.byref(ms), C.byref(n))
        return k.value, ms.value, n.value

    unfused_best, _ = _timed(ctx, world, dist, torch, once_unfused, repeats=1)
    jit0 = jit_stats()
    once()   # untimed: a call of this size compiles the plan's kernel on first use (cached per process afterwards)
    jit1 = jit_stats()
    best, _ = _timed(ctx, world, dist, torch, once, repeats=2)
    jit2 = jit_stats()
    kept_all = state["kept"]
    if world > 1:
        t = torch.tensor([float(kept_all)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        kept_all = int(t.item())
    assert state["fused"][0] == kept_all, "COUNT(*) over all groups must equal the rows the filter kept"
    assert state["fused"] == state["unfused"], "fused and unfused plans must agree bit for bit (dyadic payloads)"
    prog.close()
    for ptr in d_in + d_out + [d_count]:
        ctx.free(ptr)
    return {"metric": "rows/sec, Q1 shape: Filter(ship<=D) -> Compute(disc_price, charge) -> GroupAggregate({rf,ls}; 5xSUM, COUNT) (BASELINE config 5 shape)",
            "value": world * rows / best, "unit": "rows/s", "rows_per_gpu": rows, "groups": int(state["groups"]),
            "seconds": best, "whole_table_two_step_seconds": unfused_best, "selectivity": kept_all / float(world * rows),
            "kernels": ("ssb_group_update_program: per 64M-row slice ONE kernel, ssb_jit_rows = filter + compute + aggregation "
                        "compiled at run time for this plan (NVRTC -> sm_100a, csrc/jit_rows.h); compilation happens in the "
                        "untimed first call and is reported in jit; the two-kernel form is timed beside it")
                       if jit2[2] > jit1[2] else
                       ("ssb_group_update_program: per 64M-row slice ONE kernel, expr_sink_kernel (filter + compute + aggregation "
                        "into per-thread accumulators, interpreted); the two-kernel form is timed beside it"),
            "jit": {"kernels_compiled": jit1[0] - jit0[0], "compile_ms": round(jit1[1] - jit0[1], 1),
                    "launches_in_timed_calls": jit2[2] - jit1[2]},
            "algorithmic_gbs_per_gpu": rows * 56 / best / 1e9, "check": "sum of COUNT(*) == rows kept by the filter; sliced == whole-table two-step (COUNT and SUM(charge) bit-exact)",
            "exchange": "ssb_shard_group_merge over the 6-group partial tables" if world > 1 else "none"}


def hash_join_aux(capi, ctx, rank, world, probe_rows, build_rows, dist, torch, comm=None):
    """BASELINE config 4 shape, row-range sharded: HashJoin(INNER, fk = pk, UNIQUE) with the build
    side a permutation of [0, B) (B = world x build_rows), probe keys uniform over the build keys,
    result {fk, lv, payload}. One rank: ssb_join_build + ssb_join_probe_materialize (the probe writes the result columns). Several ranks:
    ShardedHashJoin (hash partition kernel, all-to-all over NCCL, local join, return trip)."""
    lib = ctx.lib
    I64 = capi.INT64
    total_build = world * build_rows
    mult = 1000000007
    while np.gcd(mult, total_build) != 1:
        mult += 2
    if world == 1:
        pk, pay = ctx.malloc(build_rows * 8 + 256), ctx.malloc(build_rows * 8 + 256)
        fk, lv = ctx.malloc(probe_rows * 8 + 256), ctx.malloc(probe_rows * 8 + 256)
        o_fk, o_lv, o_pay = (ctx.malloc(probe_rows * 8 + 256) for _ in range(3))
        ptrs = dict(pk=pk, pay=pay, fk=fk, lv=lv)
        keep = [pk, pay, fk, lv, o_fk, o_lv, o_pay]
    else:
        tens = {nm: torch.empty(n_, dtype=torch.int64, device="cuda")
                for nm, n_ in [("pk", build_rows), ("pay", build_rows), ("fk", probe_rows), ("lv", probe_rows)]}
        ptrs = {nm: t.data_ptr() for nm, t in tens.items()}
        torch.cuda.synchronize()
    ctx.generate(ptrs["pk"], build_rows, rank * build_rows, SEED, 40, 4, mult, total_build)
    ctx.generate(ptrs["pay"], build_rows, rank * build_rows, SEED, 41, 0, 0, 0)
    ctx.generate(ptrs["fk"], probe_rows, rank * probe_rows, SEED, 42, 1, 0, total_build)
    ctx.generate(ptrs["lv"], probe_rows, rank * probe_rows, SEED, 43, 0, 0, 0)
    ctx.sync()
    state = {}
    if world == 1:
        def once():
            j = C.c_void_p()
            ctx.check(lib.ssb_join_build(ctx.h, 1, _cols(capi, [(ptrs["pk"], None, I64)]), build_rows, 1, C.byref(j)))
            n = C.c_int64()
            # UNIQUE keys: the probe writes the three result columns itself (no row-id lists, no gathers)
            ctx.check(lib.ssb_join_probe_materialize(j, _cols(capi, [(ptrs["fk"], None, I64)]), probe_rows, 0,
                                                     2, _cols(capi, [(ptrs["fk"], None, I64), (ptrs["lv"], None, I64)]),
                                                     1, _cols(capi, [(ptrs["pay"], None, I64)]),
                                                     _cols(capi, [(o_fk, None, I64), (o_lv, None, I64), (o_pay, None, I64)]), None, C.byref(n)))
            assert n.value <= probe_rows
            ctx.sync()
            state["pairs"] = n.value
            lib.ssb_join_destroy(j)
    else:
        from supersonic_b200.distributed import CudaJoinKernels, ShardedHashJoin
        kern = CudaJoinKernels(ctx)

        def run_with(strategy):
            join = ShardedHashJoin(kern, strategy=strategy)

            def once_():
                rows_, lcols, rcols, _ = join.run([(tens["fk"], I64)], [(tens["fk"], I64), (tens["lv"], I64)],
                                                  [(tens["pk"], I64)], [(tens["pay"], I64)], join_type=0, uniqueness=1)
                state["pairs"] = int(rows_.numel())
            return once_
        # the Python-orchestrated forms, timed for the record: probe rows redistributed (hash partition + three
        # all-to-alls), the whole build side all-gathered with the whole table built on every rank, and the
        # replicated form over torch.distributed
        a2a_best, _ = _timed(ctx, world, dist, torch, run_with("all_to_all"), repeats=1)
        state["all_to_all_seconds"] = a2a_best
        bc_best, _ = _timed(ctx, world, dist, torch, run_with("broadcast"), repeats=1)
        state["broadcast_seconds"] = bc_best
        rp_best, _ = _timed(ctx, world, dist, torch, run_with("auto"), repeats=1)
        state["python_replicate_seconds"] = rp_best
        # ... and the same strategy behind the C ABI (ssb_shard_join_*: partition, ONE grouped NCCL exchange of key +
        # payload, per-part compact tables, all-gather of tables + payload, local probe; no probe row moves)
        o_fk, o_lv, o_pay = (torch.empty(probe_rows, dtype=torch.int64, device="cuda") for _ in range(3))

        def once():
            j = C.c_void_p()
            ctx.check(lib.ssb_shard_join_build(comm.h, _cols(capi, [(ptrs["pk"], None, I64)]), 1,
                                               _cols(capi, [(ptrs["pay"], None, I64)]), build_rows, C.byref(j)))
            n = C.c_int64()
            which = (C.c_int32 * 1)(0)
            ctx.check(lib.ssb_shard_join_probe_materialize(j, _cols(capi, [(ptrs["fk"], None, I64)]), probe_rows, 0,
                                                           2, _cols(capi, [(ptrs["fk"], None, I64), (ptrs["lv"], None, I64)]), 1, which,
                                                           _cols(capi, [(o_fk.data_ptr(), None, I64), (o_lv.data_ptr(), None, I64),
                                                                        (o_pay.data_ptr(), None, I64)]), None, C.byref(n)))
            assert n.value <= probe_rows
            ctx.sync()
            state["pairs"] = n.value
            state["form"] = lib.ssb_shard_join_form(j)
            lib.ssb_shard_join_destroy(j)
    best, _ = _timed(ctx, world, dist, torch, once, repeats=2)
    pairs = state["pairs"]
    if world > 1:
        t = torch.tensor([float(pairs)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        pairs = int(t.item())
        del tens
        torch.cuda.empty_cache()
    else:
        for ptr in keep:
            ctx.free(ptr)
    assert pairs == world * probe_rows, "every probe key exists exactly once in the build side"
    alg = 32.0 * build_rows + 40.0 * probe_rows
    return {"metric": "rows/sec (probe rows), HashJoin INNER UNIQUE on INT64 key, result {fk, lv, payload} (BASELINE config 4 shape)",
            "value": world * probe_rows / best, "unit": "rows/s", "probe_rows_per_gpu": probe_rows,
            "build_rows_per_gpu": build_rows, "pairs": pairs, "seconds": best,
            "algorithmic_gbs_per_gpu": alg / best / 1e9, "check": "pairs == probe rows (every fk has one pk)",
            "exchange": ((("ssb_shard_join_* (C ABI, NCCL inside libssb200.so), dense integer keys: key + payload columns "
                           "all-gathered as they are (one grouped exchange), every rank builds the direct index key - min -> row "
                           "(one scatter, no hashing, no table travels), local probe reading the payload by key; ")
                          if state.get("form") == 1 else
                          ("ssb_shard_join_* (C ABI, NCCL inside libssb200.so): build rows to the owner of their key's hash part "
                           "(one grouped send/recv), one compact table per rank, all-gather of the tables + payload, local probe "
                           "of the key's part; ")) +
                         "for the record, orchestrated from Python over torch.distributed: all-to-all form "
                         "%.4f s, broadcast form (whole table built on every rank) %.4f s, replicated form %.4f s"
                         % (state["all_to_all_seconds"], state["broadcast_seconds"], state["python_replicate_seconds"]))
            if world > 1 else "none"}


def fill_host_column(capi, name, arr, first_row, threads):
    """Fills `arr` (int64, possibly pinned) with column `name` of the synthetic table, in slices on `threads`
    threads (ctypes releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    kind, lo, span = GEN[name]
    lib = capi.load()
    rows = arr.shape[0]
    step = max(1 << 20, (rows + threads * 4 - 1) // (threads * 4))

    def part(begin):
        n = min(step, rows - begin)
        lib.ssb_generate_host(arr.ctypes.data + begin * 8, n, first_row + begin, SEED, COLS.index(name), kind, lo, span)
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(part, range(0, rows, step)))


def sort_aux(capi, ctx, rank, world, rows, dist, torch):
    """Sort: ssb_sort_permutation over one INT64 key column of uniformly random 64-bit values
    (every one of the eight 8-bit digits varies: eight one-sweep passes over (key image, row id) pairs) per rank;
    no exchange (each rank sorts its own shard; the sharded sample sort is verified by tests/test_multi_gpu_nccl.py).
    Algorithmic bytes per pass and pair: 16 read + 16 written."""
    key = ctx.malloc(rows * 8 + 256)
    perm = ctx.malloc(rows * 8 + 256)
    ctx.generate(key, rows, rank * rows, SEED, 50, 0, 0, 0)
    desc = (C.c_int32 * 1)(0)

    def once():
        ctx.check(ctx.lib.ssb_sort_permutation(ctx.h, 1, _cols(capi, [(key, None, capi.INT64)]), desc, rows, perm))
        ctx.sync()

    best, _ = _timed(ctx, world, dist, torch, once, repeats=2)
    # order check on a sample of adjacent pairs of the permutation
    k = np.empty(rows, dtype=np.int64)
    p_ = np.empty(rows, dtype=np.int64)
    ctx.d2h(k, key)
    ctx.d2h(p_, perm)
    step = max(1, rows // 1_000_000)
    sample = k[p_[::step]]
    assert np.all(np.diff(sample) >= 0), "sorted order violated"
    ctx.free(key)
    ctx.free(perm)
    passes = 8
    return {"metric": "rows/sec, Sort: stable radix sort permutation of one INT64 key column (uniform 64-bit keys)",
            "value": world * rows / best, "unit": "rows/s", "rows_per_gpu": rows, "seconds": best, "passes": passes,
            "algorithmic_gbs_per_gpu": rows * 32.0 * passes / best / 1e9,
            "check": "keys gathered through the permutation are non-decreasing (1M-row sample)"}


def host_column(capi, name, rows, first_row=0):
    kind, lo, span = GEN[name]
    out = np.empty(rows, dtype=np.int64)
    capi.load().ssb_generate_host(out.ctypes.data, rows, first_row, SEED, COLS.index(name), kind, lo, span)
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    from supersonic_b200 import capi, ssplan
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # NCCL prints its version banner (and warnings) to stdout: send them to a file, stdout carries one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/ssb200_nccl.%h.%p.log")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = capi.Context(local)
    rows = args.rows
    prog = build_program(capi, ctx)
    # --- resident inputs: this rank's row range of the global table
    first = rank * rows
    d_cols = {}
    for name in COLS:
        d_cols[name] = ctx.malloc(rows * 8 + 256)
        kind, lo, span = GEN[name]
        ctx.generate(d_cols[name], rows, first, SEED, COLS.index(name), kind, lo, span)
    d_out = ctx.malloc(rows * 8 + 256)
    d_count = ctx.malloc(8)
    ctx.sync()
    inputs = [(d_cols[c], None, capi.INT64) for c in "abcd"]
    outputs = [(d_out, None, capi.INT64)]

    def barrier():
        ctx.sync()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    ctx.enable_timing(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        prog.run(inputs, rows, outputs, d_count)
    ctx.sync()
    launches0 = ctx.launches()
    barrier()
    kernel_ms = []
    sampler.mark(0)
    ctx.timer_start()
    for _ in range(args.steps):
        prog.run(inputs, rows, outputs, d_count)
        if args.per_kernel_timing:
            kernel_ms.append(ctx.last_kernel_ms())
    total_ms = ctx.timer_stop()
    sampler.mark(1)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launches() - launches0
    kept = np.zeros(1, dtype=np.int64)
    ctx.d2h(kept, d_count)
    if not args.per_kernel_timing:
        # average launch duration of the dominant kernel from the event pair around each launch
        for _ in range(3):
            prog.run(inputs, rows, outputs, d_count)
            kernel_ms.append(ctx.last_kernel_ms())
    # max over ranks
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        k = torch.tensor([float(kept[0])], device="cuda", dtype=torch.float64)
        dist.all_reduce(k, op=dist.ReduceOp.SUM)
        kept_total = int(k.item())
    else:
        kept_total = int(kept[0])
    ms_per_step = total_ms / args.steps
    value = world * rows / (ms_per_step * 1e-3)

    result = None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        sel = kept_total / float(world * rows)
        alg_bytes = rows * (32.0 + 8.0 * sel)
        k_ms = float(np.mean(kernel_ms))
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        result = {
            "metric": METRIC,
            "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": {"workload": workload(rows), "residency": "value: inputs resident in HBM; e2e: pinned host buffers",
                       "rows_per_gpu": rows, "selectivity": sel, "l2": "inputs (%.1f GB per step) larger than L2"
                       % (rows * 32 / 1e9), "parallelism": "row-range shards, no collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic(rows), "peak_source": peak_src,
                         "kernel": "expr_kernel<96,8> (768-row tiles, 3 CTAs/SM)", "kernel_ms": k_ms,
                         "algorithmic_bytes_per_row": 32.0 + 8.0 * sel},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "kept_rows": kept_total,
        }
    # ---- sharded aggregate (config 3 shape): frees the filter columns first
    for name in COLS:
        ctx.free(d_cols[name])
    ctx.free(d_out)
    comm = None
    if world > 1:
        from supersonic_b200.distributed import make_comm
        comm = make_comm(ctx)
    aux_group = group_by_aux(capi, ctx, rank, world, min(rows, args.group_rows), dist, torch, comm)
    aux_q1 = q1_aux(capi, ctx, rank, world, min(rows, args.q1_rows), dist, torch, comm)
    aux_join = hash_join_aux(capi, ctx, rank, world, min(rows, args.join_probe_rows),
                             max(1, min(rows, args.join_probe_rows) // 10), dist, torch, comm)
    aux_sort = sort_aux(capi, ctx, rank, world, min(rows, args.sort_rows), dist, torch)
    if comm is not None:
        comm.close()
    if rank == 0:
        peak_gbs = measured_peak_gbs()[0]
        for a_ in (aux_group, aux_q1, aux_join, aux_sort):
            a_["frac_of_hbm_peak"] = a_["algorithmic_gbs_per_gpu"] / peak_gbs
        result["aux"] = {"group_by": aux_group, "q1": aux_q1, "hash_join": aux_join, "sort": aux_sort}
        # the sharded operators in one compact object (rows/s over all ranks, seconds per pass): the 1 -> N curves
        # of the configurations whose timed region contains an exchange
        result["scale_aux"] = {"unit": "rows/s",
                               "group_by_c3": aux_group["value"], "group_by_c3_s": aux_group["seconds"],
                               "q1_c5": aux_q1["value"], "q1_c5_s": aux_q1["seconds"],
                               "hash_join_c4": aux_join["value"], "hash_join_c4_s": aux_join["seconds"],
                               "sort": aux_sort["value"], "sort_s": aux_sort["seconds"],
                               "frac_of_hbm_peak": {"group_by_c3": aux_group["frac_of_hbm_peak"], "q1_c5": aux_q1["frac_of_hbm_peak"],
                                                    "hash_join_c4": aux_join["frac_of_hbm_peak"], "sort": aux_sort["frac_of_hbm_peak"]}}
    # ---- end to end through the supersonic.h mirror with pinned host buffers: the headline table
    # itself (rows per GPU x 4 read columns = 32 GB of pinned host memory per rank) when the box has
    # the memory, else the largest table that leaves half of the available RAM free
    e2e_rows = rows if args.e2e_rows <= 0 else min(args.e2e_rows, rows)
    try:
        import psutil
        budget = psutil.virtual_memory().available // (2 * max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world))))
        fit = int(budget // 32) // (1 << 22) * (1 << 22)
        if fit < e2e_rows:
            e2e_rows = max(1 << 22, fit)
    except Exception:
        pass
    host = {}
    gen_threads = max(1, min(16, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world)))))
    for name in "abcd":
        p = ctx.malloc_host(e2e_rows * 8)
        arr = np.ctypeslib.as_array((C.c_int64 * e2e_rows).from_address(p))
        fill_host_column(capi, name, arr, first, gen_threads)
        host[name] = (p, arr)
    plan_lib = ssplan.PlanLib(os.path.join(ROOT, "supersonic_b200", "lib", "libssb200_plan.so"))
    cols = [ssplan.Column(nm, ssplan.INT64, host[nm][1]) for nm in "abcd"]
    for c, nm in zip(cols, "abcd"):
        c.data = host[nm][1]            # keep the pinned buffer (no numpy copy)
    e2e_steps = max(1, args.steps if args.e2e_steps <= 0 else args.e2e_steps)
    e2e_kept = 0
    moved0 = (C.c_uint64(), C.c_uint64())

    def e2e_step():
        r = plan_lib.run(PLAN, [cols], next_max_rows=1 << 22, flags=ssplan.SSPLAN_DISCARD)
        if r.code != 0:
            raise RuntimeError("e2e plan failed: %d %s" % (r.code, r.error))
        return r.rows

    e2e_step()                           # one untimed pass (program compilation, lane buffers)
    capi.load().ssb_transfer_bytes(C.byref(moved0[0]), C.byref(moved0[1]))
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_kept = e2e_step()            # every step: H2D of the inputs, kernel, D2H of the kept rows, drained to the end
    barrier()
    e2e_t = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_t], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_t = float(t.item())
    moved1 = (C.c_uint64(), C.c_uint64())
    capi.load().ssb_transfer_bytes(C.byref(moved1[0]), C.byref(moved1[1]))
    h2d_step = (moved1[0].value - moved0[0].value) // e2e_steps
    d2h_step = (moved1[1].value - moved0[1].value) // e2e_steps
    for name in "abcd":
        ctx.free_host(host[name][0])
    if rank == 0:
        result["e2e"] = {"value": world * e2e_rows / e2e_t, "unit": "rows/s", "rows_per_gpu": e2e_rows, "steps": e2e_steps,
                         "kept_rows_per_gpu": int(e2e_kept), "seconds_per_step": e2e_t,
                         "h2d_bytes_per_step": int(h2d_step), "d2h_bytes_per_step": int(d2h_step),
                         "host_bytes_per_step": int(e2e_rows * 32),
                         "api": "supersonic::Filter/Compute/ScanView cursors via the plan driver, pinned host views",
                         "transfer": "64-bit integer columns whose chunk fits 32 bits cross PCIe as 32-bit values "
                                     "(host threads narrow and verify every value, the kernel widens them: lossless); "
                                     "only the kept rows come back; h2d/d2h bytes are the copies actually issued"}
        # ---- CPU baseline: the reference itself, one thread, bounded sample
        result["cpu_baseline"] = cpu_reference_sample(args.cpu_rows, threads=1)
        emit(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


_M64 = (1 << 64) - 1


def np_column(name, rows, first_row=0):
    """The same synthetic column as ssb_generate / ssb_generate_host (kind 0), in numpy: the reference arm
    loads nothing of this repo's native code."""
    kind, lo, span = GEN[name]
    assert kind == 0
    with np.errstate(over="ignore"):
        x = np.arange(first_row, first_row + rows, dtype=np.uint64)
        x += np.uint64((SEED ^ ((COLS.index(name) * 0x9E3779B97F4A7C15) & _M64)) & _M64)
        x += np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
        if span:
            x &= np.uint64(span - 1)
        x += np.uint64(lo & _M64)
    return x.view(np.int64)


def _ref_worker(arg):
    rows, first = arg
    from supersonic_b200 import ssplan
    ref = ssplan.PlanLib(os.path.join(ROOT, "oracle", "_ref", "libssref.so"))
    cols = [ssplan.Column(nm, ssplan.INT64, np_column(nm, rows, first)) for nm in "abcd"]
    best = None
    for _ in range(3):
        r = ref.run(PLAN, [cols], next_max_rows=1024, flags=ssplan.SSPLAN_DISCARD)
        if r.code != 0:
            raise RuntimeError(r.error)
        best = r.drain_seconds if best is None else min(best, r.drain_seconds)
    return rows, best, r.rows


def cpu_reference_sample(rows, threads):
    """Times the unmodified reference (oracle/_ref/libssref.so) over `rows` rows per thread."""
    import multiprocessing as mp
    ref_path = os.path.join(ROOT, "oracle", "_ref", "libssref.so")
    if not os.path.exists(ref_path):
        return {"value": None, "unit": "rows/s", "cores": threads, "kind": "reference",
                "sample": "oracle/_ref/libssref.so missing"}
    t0 = time.perf_counter()
    if threads == 1:
        res = [_ref_worker((rows, 0))]
        wall = res[0][1]
    else:
        with mp.get_context("fork").Pool(threads) as pool:
            res = pool.map(_ref_worker, [(rows, i * rows) for i in range(threads)])
        wall = max(r[1] for r in res)
    total = sum(r[0] for r in res)
    return {"value": total / wall, "unit": "rows/s", "cores": threads, "kind": "reference",
            "sample": "%d rows per thread of the same plan, Next(1024) drain, best of 3, wall %.2fs"
                      % (rows, time.perf_counter() - t0)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    args.cpu_rows = max(1 << 20, min(args.cpu_rows, 400_000_000 // cores))   # bounded memory and time
    values = []
    base = None
    for _ in range(max(1, args.steps)):
        base = cpu_reference_sample(args.cpu_rows, threads=cores)
        values.append(base["value"])
    v = float(np.mean(values))
    out = {"impl": "reference", "metric": METRIC,
           "value": v, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": cores * args.cpu_rows / v * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "int64", "data": "synthetic",
           "config": {"workload": workload(args.rows), "rows_per_gpu": args.rows,
                      "sample": "the reference's own CPU cursors over a bounded sample of that table: %d independent "
                                "single-thread cursors (the engine is single-threaded), %d rows each, Next(1024) drain"
                                % (cores, args.cpu_rows)},
           "cpu_baseline": dict(base, value=v),
           "e2e": {"value": v, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(json.dumps(out))


_RESULT_STREAM = None


def claim_stdout():
    """stdout carries exactly one JSON line: everything else that libraries print there (NCCL's
    version banner, for one) is sent to stderr by pointing fd 1 at fd 2 for the run."""
    global _RESULT_STREAM
    sys.stdout.flush()
    _RESULT_STREAM = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _RESULT_STREAM or sys.stdout
    out.write(line + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000_000, help="rows per GPU (BASELINE C2: 1e9)")
    ap.add_argument("--e2e-rows", type=int, default=0, help="rows per GPU of the end-to-end leg (0 = --rows, bounded by host RAM)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed steps of the end-to-end leg (0 = --steps)")
    ap.add_argument("--cpu-rows", type=int, default=20_000_000)
    ap.add_argument("--group-rows", type=int, default=1_000_000_000, help="rows per GPU of the aux group-by")
    ap.add_argument("--q1-rows", type=int, default=600_000_000, help="rows per GPU of the aux Q1-shape plan (C5: 6e8)")
    ap.add_argument("--join-probe-rows", type=int, default=125_000_000,
                    help="probe rows per GPU of the aux hash join (C4: 1e9 over 8 GPUs); build side = a tenth")
    ap.add_argument("--sort-rows", type=int, default=200_000_000, help="rows per GPU of the aux sort")
    ap.add_argument("--per-kernel-timing", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
