#!/usr/bin/env python
"""bench_configs.py — the five BASELINE.json workloads (one GPU's share of each) through the C ABI, one JSON line each.

Not the driver's contract (that is bench.py, configs[1]); this is the measurement of the other SURVEY.md §8d
configs: rows/s with segments resident in HBM, CUDA-event kernel times, algorithmic bytes and a parity spot-check of
segment 0 against the oracle (full size).

    python bench_configs.py [--only 1,3,4,5] [--steps 10]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="1,2,3,4,5")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--scale", type=float, default=1.0, help="shrink rows per segment (smoke runs)")
    args = ap.parse_args()
    only = {int(x) for x in args.only.split(",")}

    import torch
    from oracle import oracle
    from pinot_b200 import datagen, native
    from pinot_b200.query import parse_sql
    from tests.parity import assert_rows_equal, oracle_rows

    torch.cuda.set_device(0)
    native.init(0)
    peak = 6564.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass

    def rows_per(n):
        return max(2048, int(n * args.scale))

    configs = []
    if 1 in only:
        configs.append(dict(id=1, n_seg=1, rows=rows_per(1_000_000), cols=["c0", "c1"],
                            sql=lambda segs: datagen.config1_sql(segs[0]), exact=True,
                            name="1 segment x 1M rows, SELECT SUM(c0) WHERE c1 > k (50 %)"))
    if 2 in only:
        configs.append(dict(id=2, n_seg=8, rows=rows_per(12_500_000), cols=datagen.CONFIG2_COLUMNS,
                            sql=lambda segs: datagen.config2_sql(segs, 16), exact=True,
                            name="8 x 12.5M rows, c1 IN(16) AND c2<k GROUP BY d0,d1,d2, 4 aggs"))
    if 3 in only:
        def sql3(segs):
            d1 = segs[0].columns["c1"].dictionary_values()
            d3 = segs[0].columns["c3"].dictionary_values()
            d0 = segs[0].columns["d0"].dictionary_values()
            return (f"SELECT d0, d1, d2, d3, d4, SUM(m0), COUNT(*), MIN(m1), MAX(m2) FROM t WHERE (c1 IN ({', '.join(str(int(v)) for v in d1[::125][:8])}) "
                    f"OR c3 = {int(d3[5])}) AND d0 IN ({int(d0[1])}, {int(d0[6])}) GROUP BY d0, d1, d2, d3, d4 LIMIT 100000")
        configs.append(dict(id=3, n_seg=8, rows=rows_per(1_562_500), cols=["c1", "c3", "d0", "d1", "d2", "d3", "d4", "m0", "m1", "m2"],
                            sql=sql3, exact=True,
                            name="8 x 1.5625M rows (one GPU's share of 64 segments), inverted-index (c1 IN(8) OR c3=v) AND d0 IN(2), GROUP BY 5 dims"))
    if 4 in only:
        def sql4(segs):
            d2 = segs[0].columns["c2"].dictionary_values()
            return f"SELECT s0, DISTINCTCOUNT(c0), SUM(x0) FROM t WHERE c2 < {int(d2[len(d2) // 2])} GROUP BY s0 LIMIT 100000"
        configs.append(dict(id=4, n_seg=8, rows=rows_per(15_625_000), cols=["s0", "c0", "c2", "x0"], sql=sql4, exact=False,
                            name="8 x 15.625M rows (one GPU's share of 64), GROUP BY s0 (STRING dict, 10k groups) DISTINCTCOUNT(c0) + SUM(x0 raw DOUBLE) WHERE c2<k"))
    if 5 in only:
        configs.append(dict(id=5, n_seg=8, rows=rows_per(12_500_000), cols=["k0", "m0"],
                            sql=lambda segs: "SET numGroupsLimit = 20000000; SELECT k0, SUM(m0), COUNT(*) FROM t GROUP BY k0 LIMIT 100000000",
                            exact=True, name="8 x 12.5M rows (one GPU's share of 32), GROUP BY k0 (raw LONG, ~10M groups) SUM(m0), COUNT(*)"))

    for cfg in configs:
        t0 = time.time()
        segs = datagen.make_table(cfg["n_seg"], cfg["rows"], columns=cfg["cols"])
        q = parse_sql(cfg["sql"](segs))
        gen_s = time.time() - t0
        staged = [native.StagedSegment(s) for s in segs]
        group = native.SegmentGroup(staged)
        prepared = native.prepare(q)
        rows_total = sum(s.num_docs for s in segs)
        # algorithmic bytes: touched columns x stored bits; bitmap operands count numDocs/8 each
        _, preds = q.filter_postfix()
        touched = set(q.group_by) | {a.column for a in q.aggregations if a.column}
        n_bitmaps = 0
        for p in preds:
            c = segs[0].columns[p.column]
            if c.inverted_index is not None and p.column not in q.skip_indexes and int(p.type) != 4:
                n_bitmaps += max(1, len(p.values))      # one flat bitmap operand per dictId (numDocs/8 bytes each)
            else:
                touched.add(p.column)                   # scanned predicate column
        alg = 0.0
        for s in segs:
            bits = sum((s.columns[c].bits_per_element if s.columns[c].has_dictionary else 8 * s.columns[c].dict_entry_bytes) for c in touched)
            alg += s.num_docs * (bits + n_bitmaps) / 8.0
        for _ in range(3):
            native.execute(group, q, native.PB_Q_COMBINE, prepared).free()
        walls, kern, filt, agg = [], [], [], []
        torch.cuda.synchronize()
        for _ in range(args.steps):
            ts = time.perf_counter()
            r = native.execute(group, q, native.PB_Q_COMBINE, prepared)
            walls.append(time.perf_counter() - ts)
            kern.append(r.scan_kernel_ms)
            f_, a_ = r.phase_ms()
            filt.append(f_); agg.append(a_)
            ng = r.tables[0].num_groups
            matched = r.tables[0].stats["num_docs_scanned"]
            launches = r.kernel_launches
            r.free()
        # parity spot check: segment 0 alone, device vs oracle
        g0 = native.SegmentGroup([staged[0]])
        r0 = native.execute(g0, q)
        o0 = oracle.execute(segs[0], q)
        parity = "ok"
        try:
            assert_rows_equal(r0.tables[0].rows(), oracle_rows(o0), q, exact_float=cfg["exact"], what=f"config {cfg['id']} segment 0")
            assert r0.tables[0].stats["num_docs_scanned"] == o0.stats["num_docs_scanned"]
        except AssertionError as e:
            parity = "MISMATCH: " + str(e)[:300]
        r0.free()
        g0.release()
        wall = float(np.median(walls))
        k_ms = float(np.mean(kern))
        line = {"config": cfg["id"], "workload": cfg["name"], "rows": rows_total, "groups": int(ng), "docs_matched": int(matched),
                "rows_per_s_wall": rows_total / wall, "wall_ms_median": 1000 * wall,
                "kernel_ms": k_ms, "filter_kernel_ms": float(np.mean(filt)), "agg_kernel_ms": float(np.mean(agg)),
                "rows_per_s_kernels": rows_total / (k_ms * 1e-3), "launches": int(launches),
                "algorithmic_bytes": alg, "frac_of_measured_hbm_kernels": alg / (k_ms * 1e-3) / 1e9 / peak,
                "parity_segment0_vs_oracle": parity, "gen_s": round(gen_s, 1), "sql": cfg["sql"](segs)[:200]}
        print(json.dumps(line), flush=True)
        group.release()
        for s in staged:
            s.release()
        del segs, staged
    return 0


if __name__ == "__main__":
    sys.exit(main())
