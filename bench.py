#!/usr/bin/env python
"""bench.py — scanned rows/s of the per-segment filter + group-by hot path (BASELINE.json metric).

Workload at N=1 (BASELINE.json configs[1]): 8 segments x 12.5 M rows = 100 M rows of the 20-column synthetic table
(only the 8 columns the query touches are materialised), inverted index on c1 disabled so both predicates scan:

    SELECT d0,d1,d2, SUM(m0), COUNT(*), MIN(m1), MAX(m2) FROM t
    WHERE c1 IN (16 values) AND c2 < k(50 %) GROUP BY d0,d1,d2            -- 4 096 groups, ~0.8 % selectivity

A step = one pass of the whole query over all segments of this rank through the C ABI (host planning layer ->
pb_query_execute, merged result table back in pinned host memory).  `value` = rows / wall time of K steps with the
segments already resident in HBM; `e2e` = the same call sequence starting from page-locked HOST buffers inside the timed
region (pb_segment_stage + execute + result read-back; headline policy PB_Q_GATHER_IN_PLACE = copy the predicate columns,
gather the group-by / aggregation columns of the matching rows over PCIe; the copy-every-touched-column policy is
measured alongside as e2e.legs.stage_all).  N > 1: every rank owns its own 8
segments (weak scaling), per-rank dense tables are merged over NCCL (one all-gather of the table block + a merge kernel;
PB_MERGE=allreduce selects three in-place all-reduces instead), rank 0 finalises.

`--impl reference` times the CPU restatement of the reference path (oracle/, the one place this file may run it
besides the cpu_baseline leg) on the host cores, one thread per segment.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "scanned rows/sec for filter+groupby(3 dims,4 aggs) @1/2/4/8 B200; %HBM BW"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--segments", type=int, default=8)
    ap.add_argument("--docs-per-segment", type=int, default=12_500_000)
    ap.add_argument("--in-values", type=int, default=16)
    ap.add_argument("--flags", type=int, default=0, help="extra PB_Q_* flags (A/B: 4 = generic predicate path, 8 = no TMA)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-policy", choices=("in_place", "stage_all"), default="in_place",
                    help="cold-segment staging policy of the e2e leg (stage_all is always measured and reported alongside)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML in a thread; same fields as the
    nvidia-smi line of B200_PROFILING.md, without forking a process inside the timed region)."""

    def __init__(self, gpu_index: int, period_s: float = 0.001):
        self.gpu, self.period = gpu_index, period_s
        self.sm, self.reasons, self.smmax = [], set(), None
        self._stop = threading.Event()
        self.t = None
        self.err = None

    def mark(self):
        """samples taken before this call (warm-up) are dropped"""
        self.sm = []
        self.reasons = set()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.smmax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:   # pragma: no cover
            self.err = repr(e)
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, b in bits.items():
                    if r & b:
                        self.reasons.add(k)
            except Exception as e:   # pragma: no cover
                self.err = repr(e)
                break
            self._stop.wait(self.period)

    def stop(self):
        self._stop.set()
        if self.t:
            self.t.join(timeout=1)
        out = {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smmax,
               "reasons": sorted(self.reasons), "samples": len(self.sm)}
        if self.err:
            out["error"] = self.err
        return out


def build_table(args, rank):
    from pinot_b200 import datagen
    segs = datagen.make_table(args.segments, args.docs_per_segment, columns=datagen.CONFIG2_COLUMNS,
                              first_index=rank * args.segments)
    return segs


def algorithmic_bytes(segs, q, part="all"):
    """sum over segments of numDocs x sum over touched columns of storedBits / 8 (full-scan convention).
    part = "filter": only the predicate columns (what pb_filter_kernel streams); "agg": the group-by/metric columns."""
    _, preds = q.filter_postfix()
    fcols = {p.column for p in preds}
    acols = set(q.group_by) | {a.column for a in q.aggregations if a.column}
    touched = fcols | acols if part == "all" else (fcols if part == "filter" else acols - fcols)
    total = 0.0
    for s in segs:
        bits = 0
        for c in touched:
            col = s.columns[c]
            bits += col.bits_per_element if col.has_dictionary else 8 * col.dict_entry_bytes
        total += s.num_docs * bits / 8.0
    return total


def oracle_query_all_threads(segs, q, threads):
    """One CombineOperator-style pass on the CPU: one task per segment on `threads` threads, then the merge."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(lambda s: oracle.execute(s, q), segs))
    return oracle.combine_numeric(res)


def run_reference(args, rank, world):
    """The reference's own CPU path (restated in C: oracle/) on the host cores."""
    if rank != 0:
        return
    from oracle import oracle
    from pinot_b200 import datagen
    from pinot_b200.query import parse_sql
    oracle.build()
    segs = build_table(args, 0)
    q = parse_sql(datagen.config2_sql(segs, args.in_values))
    threads = min(len(segs), os.cpu_count() or 1)
    rows = sum(s.num_docs for s in segs)
    for _ in range(args.warmup):
        oracle_query_all_threads(segs, q, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_query_all_threads(segs, q, threads)
    dt = time.perf_counter() - t0
    value = rows * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, segs),
            "cpu_baseline": {"value": value, "unit": "rows/s", "cores": threads, "kind": "port",
                             "sample": f"full {rows} row query per step, one thread per segment ({threads} threads), {args.steps} steps"},
            "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def workload_config(args, segs):
    return {"workload": f"BASELINE.json configs[1]: {len(segs)} segments x {segs[0].num_docs} rows, "
                        f"WHERE c1 IN({args.in_values}) AND c2<k GROUP BY d0,d1,d2 SUM/COUNT/MIN/MAX, skipIndexes c1=inverted",
            "segments_per_gpu": len(segs), "rows_per_gpu": sum(s.num_docs for s in segs),
            "columns_materialised": "8 touched of the 20-column table", "l2_policy": "inputs (>1 GB/GPU) larger than the 126 MB L2",
            "parallelism": f"segments sharded over {args.gpus} GPU(s); dense group tables merged over NCCL ({os.environ.get('PB_MERGE', 'allgather')})" if args.gpus > 1 else "1 GPU"}


_JSON_OUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries loaded later write banners to file descriptor 1 from C (NCCL prints
    its version there): keep a private duplicate of the real stdout for the JSON line and point fd 1 at stderr for everyone else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line: dict):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse_args()
    _claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return 0

    import torch
    import torch.distributed as dist
    from pinot_b200 import datagen, native
    from pinot_b200.query import AggOp, parse_sql

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep NCCL's banner off stdout (one JSON line only)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    native.init(local_rank)

    segs = build_table(args, rank)
    # table-wide literals must agree on every rank: dimension dictionaries are table-wide (datagen), so they do
    q = parse_sql(datagen.config2_sql(segs, args.in_values))
    rows_rank = sum(s.num_docs for s in segs)
    flags = native.PB_Q_COMBINE | args.flags

    # page-lock the host copies of the touched columns (what a server does once for its mmap'd segments)
    for s in segs:
        for c in s.columns.values():
            native.host_register(c.forward_index)

    staged = [native.StagedSegment(s) for s in segs]
    group = native.SegmentGroup(staged)

    if world > 1:
        # agree on the global dictionaries of the group-by columns (dense tables must line up across ranks)
        from pinot_b200.distributed import agree_global_dictionaries, all_gather_merge_tables, all_reduce_tables
        merge_mode = os.environ.get("PB_MERGE", "allgather")
        agree_global_dictionaries(group, q.group_by, [int(segs[0].columns[c].data_type) for c in q.group_by], dist)

    def barrier():
        if world > 1:
            dist.barrier()

    prepared = native.prepare(q)

    def step(g=None, extra_flags=0):
        """one pass of the hot path over this rank's segments; returns the Result"""
        g = g or group
        if world == 1:
            return native.execute(g, q, flags | extra_flags, prepared)
        r = native.execute(g, q, flags | extra_flags | native.PB_Q_DEFER_FINALIZE, prepared)
        if merge_mode == "allgather":
            all_gather_merge_tables(r, dist, torch)     # ONE collective on the call's stream + a device-side merge kernel
        else:
            all_reduce_tables(r, q, dist, torch)        # three small all-reduces on the call's stream
        if rank == 0:
            r.finalize()
        return r

    # ---- warm-up (the clock sampler and GC state are set up before it so nothing new starts inside the timed region) ----
    import gc
    sampler = ClockSampler(local_rank)
    sampler.start()
    gc.collect()
    gc.disable()          # no cyclic-GC pauses inside the timed region (re-enabled right after)
    for _ in range(max(args.warmup, 3)):
        r = step()
        r.free()

    # ---- timed: K steps, barrier + synchronize on both sides, max over ranks ----
    scan_ms, device_ms, launches, host_us, filt_ms, agg_ms, step_wall = [], [], 0, [], [], [], []
    barrier()
    torch.cuda.synchronize()
    sampler.mark()
    t0 = time.perf_counter()
    last = None
    for it in range(args.steps):
        ts = time.perf_counter()
        if last is not None:
            last.free()            # the operator frees a result before it runs the next query (same as the warm-up)
        r = step()
        step_wall.append(1000 * (time.perf_counter() - ts))
        scan_ms.append(r.scan_ms())
        f_, a_ = r.phase_ms()
        filt_ms.append(f_)
        agg_ms.append(a_)
        if world == 1 or rank == 0:
            launches += lib_launches(r)
            device_ms.append(getattr(r, "device_ms", 0.0))
            host_us.append(r.host_timing_us())
        last = r
    torch.cuda.synchronize()
    barrier()
    elapsed = time.perf_counter() - t0
    gc.enable()
    in_region = len(sampler.sm)
    if world > 1:
        t = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    # A K-step region of this workload lasts only a few ms, shorter than a handful of NVML reads: keep the SAME steps
    # running (untimed, same count on every rank) right after it so the clock median is taken under the identical load.
    n_extra = 0 if elapsed >= 0.25 else min(5000, int(0.25 / max(elapsed / max(args.steps, 1), 1e-5)))
    for _ in range(n_extra):
        last.free()
        last = step()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    clocks["samples_in_timed_region"] = in_region
    clocks["window"] = "timed region" if n_extra == 0 else f"timed region + {n_extra} identical untimed steps run back-to-back after it"
    if world > 1:
        sk = torch.tensor([float(np.mean(scan_ms))], dtype=torch.float64, device="cuda")
        dist.all_reduce(sk, op=dist.ReduceOp.MAX)
        scan_mean = float(sk.item())
    else:
        scan_mean = float(np.mean(scan_ms))
    rows_total = rows_rank * world
    value = rows_total * args.steps / elapsed

    num_groups = last.tables[0].num_groups if (rank == 0 and last.tables) else 0
    docs_matched = last.tables[0].stats["num_docs_scanned"] if (rank == 0 and last.tables) else 0
    d2h_bytes = 0
    if rank == 0 and last.tables:
        t0_ = last.tables[0]
        d2h_bytes = int(t0_.num_groups * (16 + 16 * len(q.aggregations)) + 64)
    last.free()

    # ---- e2e: same call sequence from HOST buffers (stage + execute + read-back) ----
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(2, min(args.steps, 5))
        docs_matched_rank = docs_matched // max(world, 1)       # merged statistics are totals over ranks; per-rank data is iid

        def e2e_leg(in_place):
            """stage from the page-locked host buffers + execute + read the result back, K times.  in_place: only the
            predicate columns are copied to HBM; the group-by / aggregation columns are gathered over PCIe from the
            mapped host buffers for the matching rows (PB_Q_GATHER_IN_PLACE)."""
            staged_bytes, in_place_cols = 0, 0

            def e2e_step():
                nonlocal staged_bytes, in_place_cols
                st = [native.StagedSegment(s) for s in segs]
                g2 = native.SegmentGroup(st)
                if world > 1:
                    for col in q.group_by:
                        g2.set_global_dictionary(col, group.export_dictionary(col))
                r2 = step(g2, native.PB_Q_GATHER_IN_PLACE if in_place else 0)
                if rank == 0:
                    _ = r2.tables[0].num_groups if r2.tables else 0      # result read-back
                staged_bytes = sum(s.device_bytes() for s in st)
                in_place_cols = r2.in_place_columns
                r2.free()
                g2.release()
                for s in st:
                    s.release()

            e2e_step()   # warm
            barrier()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            torch.cuda.synchronize()
            barrier()
            el = time.perf_counter() - t1
            if world > 1:
                t = torch.tensor([el], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                el = float(t.item())
            # in place: every gathered value is one or two 4-byte words; PCIe moves them as 32-byte sectors
            gather_values = docs_matched_rank * in_place_cols // max(len(segs), 1)     # matching rows x columns read in place
            h2d = int(staged_bytes + 32 * gather_values)
            return {"value": rows_total * e2e_steps / el, "ms_per_step": 1000 * el / e2e_steps, "h2d_bytes_per_step": h2d,
                    "staged_bytes_per_step": int(staged_bytes), "in_place_columns": int(in_place_cols)}

        legs = {"stage_all": e2e_leg(False)}
        if args.e2e_policy == "in_place":
            legs["gather_in_place"] = e2e_leg(True)
        head = legs["gather_in_place"] if "gather_in_place" in legs else legs["stage_all"]
        e2e = {"value": head["value"], "unit": "rows/s", "h2d_bytes_per_step": head["h2d_bytes_per_step"],
               "d2h_bytes_per_step": int(d2h_bytes), "steps": e2e_steps, "ms_per_step": head["ms_per_step"],
               "policy": "gather_in_place" if "gather_in_place" in legs else "stage_all", "legs": legs,
               "note": "per step: pb_segment_stage from page-locked host buffers + execute + result read-back.  stage_all copies every "
                       "touched column to HBM; gather_in_place (PB_Q_GATHER_IN_PLACE) copies the predicate columns and gathers the "
                       "group-by/aggregation columns of the matching rows over PCIe from the mapped host buffers "
                       "(h2d = staged bytes + 32-byte sectors x gathered values)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (pb_filter_kernel), CUDA events on its launching stream ----
    peak, peak_src = measured_peak_gbs()
    f_mean, a_mean = float(np.mean(filt_ms)), float(np.mean(agg_ms))
    alg_filter = algorithmic_bytes(segs, q, "filter")
    alg_all = algorithmic_bytes(segs, q, "all")
    achieved = alg_filter / (f_mean * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "scan_kernel_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("pb_filter_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "pb_filter_kernel", "kernel_ms": f_mean, "algorithmic_bytes_per_launch": alg_filter,
                "peak_source": peak_src,
                "traffic_note": "DRAM bytes of one launch from ncu (profiles/scan_kernel_traffic.json); below the algorithmic bytes because "
                                "later predicates of a selective conjunction are evaluated on the surviving rows only, their columns are not streamed",
                "whole_query": {"kernels": "pb_filter_kernel + pb_agg_kernel", "ms": f_mean + a_mean,
                                "algorithmic_bytes": alg_all, "achieved": alg_all / ((f_mean + a_mean) * 1e-3) / 1e9,
                                "frac": alg_all / ((f_mean + a_mean) * 1e-3) / 1e9 / peak,
                                "note": "BASELINE.md full-scan convention (87 bits/row); pb_agg_kernel only touches the sectors of "
                                        "rows that pass the filter, so this fraction can exceed 1 at low selectivity"},
                "agg_kernel_ms": a_mean}

    # ---- CPU baseline: the oracle (port of the reference algorithm) on this box's host cores ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        oracle.build()
        threads = min(len(segs), os.cpu_count() or 1)
        oracle_query_all_threads(segs[:1], q, 1)   # warm
        reps, t2 = 0, time.perf_counter()
        while True:
            oracle_query_all_threads(segs, q, threads)
            reps += 1
            if time.perf_counter() - t2 > 10.0 or reps >= 200:
                break
        cdt = time.perf_counter() - t2
        cpu = {"value": rows_rank * reps / cdt, "unit": "rows/s", "cores": threads, "kind": "port",
               "sample": f"{reps} passes of the full {rows_rank}-row query, one thread per segment ({threads} threads of {os.cpu_count()} cores), {cdt:.1f} s"}

    line = {"metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1000 * elapsed / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, segs),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "device_ms_per_step": float(np.mean(device_ms)) if device_ms else None,
            "scan_kernel_ms": scan_mean, "filter_kernel_ms": f_mean, "agg_kernel_ms": a_mean,
            "step_wall_ms": {"min": float(np.min(step_wall)), "median": float(np.median(step_wall)), "max": float(np.max(step_wall)),
                             "all": [round(float(x), 3) for x in step_wall]}, "host_us_by_phase": [round(float(x), 1) for x in np.mean(np.array(host_us), axis=0)] if host_us else None, "num_groups": int(num_groups), "docs_matched": int(docs_matched),
            "kernel_variant": {0: "tma+width-specialised", 4: "tma+generic", 8: "ldg+width-specialised", 12: "ldg+generic"}.get(args.flags & 12)}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def lib_launches(r):
    from pinot_b200 import native
    return native.lib().pb_result_kernel_launches(r._rh)


if __name__ == "__main__":
    sys.exit(main())
