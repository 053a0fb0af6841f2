#!/usr/bin/env python
"""bench_configs.py — the five BASELINE.json workloads through the C ABI, one JSON line each.

Not the driver's contract (that is bench.py, configs[1]); this measures the other SURVEY.md §8d configs at their stated
GPU counts: `python bench_configs.py --only 3` runs one GPU's share, and under torchrun (one process per GPU)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench_configs.py --only 3,4
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 ... bench_configs.py --only 5

the table's segments are sharded over the ranks and the per-rank tables are merged inside libpinot_b200.so
(PB_Q_ALL_RANKS: all-gather + merge kernel for dense tables and DISTINCTCOUNT bitsets, hash-partitioned all-to-all for
hash tables).  Reported: rows/s over wall time of K steps (max over ranks), CUDA-event kernel times, the NCCL merge time,
algorithmic bytes, and a parity check of the MERGED result against the oracle over all ranks' segments (DISTINCTCOUNT
value sets are only compared when the table is small enough to exchange them; see `parity`).
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
    ap.add_argument("--gen-workers", type=int, default=8)
    args = ap.parse_args()
    only = {int(x) for x in args.only.split(",")}
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import torch
    import torch.distributed as dist
    from oracle import oracle
    from pinot_b200 import datagen, native
    from pinot_b200.distributed import TorchExchange, agree_global_dictionaries, dictionary_columns, init_comm, shard_segments
    from pinot_b200.query import AggOp, parse_sql
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tests.parity import assert_rows_equal, combined_rows

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.cuda.set_device(local_rank)
    native.init(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        init_comm(TorchExchange(dist))
    oracle.build()
    peak = 6564.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass

    def rows_per(n):
        return max(2048, int(n * args.scale))

    # n_seg = segments of the WHOLE table (BASELINE.json); a single process takes one GPU's share of `gpus`
    configs = []
    if 1 in only:
        configs.append(dict(id=1, n_seg=1, gpus=1, rows=rows_per(1_000_000), cols=["c0", "c1"],
                            sql=lambda segs: datagen.config1_sql(segs[0]), exact=True,
                            name="1 segment x 1M rows, SELECT SUM(c0) WHERE c1 > k (50 %)"))
    if 2 in only:
        configs.append(dict(id=2, n_seg=8, gpus=1, rows=rows_per(12_500_000), cols=datagen.CONFIG2_COLUMNS,
                            sql=lambda segs: datagen.config2_sql(segs, 16), exact=True,
                            name="8 x 12.5M rows, c1 IN(16) AND c2<k GROUP BY d0,d1,d2, 4 aggs"))
    if 3 in only:
        def sql3(segs):
            d1 = segs[0].columns["c1"].dictionary_values()
            d3 = segs[0].columns["c3"].dictionary_values()
            d0 = segs[0].columns["d0"].dictionary_values()
            return (f"SELECT d0, d1, d2, d3, d4, SUM(m0), COUNT(*), MIN(m1), MAX(m2) FROM t WHERE (c1 IN ({', '.join(str(int(v)) for v in d1[::125][:8])}) "
                    f"OR c3 = {int(d3[5])}) AND d0 IN ({int(d0[1])}, {int(d0[6])}) GROUP BY d0, d1, d2, d3, d4 LIMIT 100000")
        configs.append(dict(id=3, n_seg=64, gpus=8, rows=rows_per(1_562_500), cols=["c1", "c3", "d0", "d1", "d2", "d3", "d4", "m0", "m1", "m2"],
                            sql=sql3, exact=True,
                            name="100M rows as 64 x 1.5625M, inverted-index (c1 IN(8) OR c3=v) AND d0 IN(2), GROUP BY 5 dims, 8 segments/GPU on 8 GPUs"))
    if 4 in only:
        def sql4(segs):
            d2 = segs[0].columns["c2"].dictionary_values()
            return f"SELECT s0, DISTINCTCOUNT(c0), SUM(x0) FROM t WHERE c2 < {int(d2[len(d2) // 2])} GROUP BY s0 LIMIT 100000"
        configs.append(dict(id=4, n_seg=64, gpus=8, rows=rows_per(15_625_000), cols=["s0", "c0", "c2", "x0"], sql=sql4, exact=False,
                            name="1B rows as 64 x 15.625M, GROUP BY s0 (STRING dict, 10k groups) DISTINCTCOUNT(c0) + SUM(x0 raw DOUBLE) WHERE c2<k, 8 GPUs"))
    if 5 in only:
        configs.append(dict(id=5, n_seg=32, gpus=4, rows=rows_per(12_500_000), cols=["k0", "m0"],
                            sql=lambda segs: "SET numGroupsLimit = 20000000; SELECT k0, SUM(m0), COUNT(*) FROM t GROUP BY k0 LIMIT 100000000",
                            exact=True, name="400M rows as 32 x 12.5M, GROUP BY k0 (raw LONG, ~10M groups) SUM(m0), COUNT(*), 4 GPUs"))

    for cfg in configs:
        t0 = time.time()
        gpus = world if world > 1 else cfg["gpus"]
        mine = shard_segments(cfg["n_seg"], rank, gpus)            # a single process runs rank 0's share of cfg["gpus"]
        segs = datagen.make_table_parallel(len(mine), cfg["rows"], columns=cfg["cols"], indices=mine,
                                           workers=args.gen_workers if cfg["rows"] >= 2_000_000 else 1)
        q = parse_sql(cfg["sql"](segs))
        gen_s = time.time() - t0
        staged = [native.StagedSegment(s) for s in segs]
        group = native.SegmentGroup(staged)
        if world > 1:
            cols = dictionary_columns(q, segs[0])
            agree_global_dictionaries(group, cols, [int(segs[0].columns[c].data_type) for c in cols], TorchExchange(dist))
        prepared = native.prepare(q)
        flags = native.PB_Q_COMBINE | (native.PB_Q_ALL_RANKS if world > 1 else 0)
        rows_rank = sum(s.num_docs for s in segs)
        rows_total = rows_rank * world
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
        alg *= world

        # ---- parity of the merged result against the oracle over ALL ranks' segments ----
        has_dc = any(a.op == AggOp.DISTINCTCOUNT for a in q.aggregations)
        hash_partitioned = world > 1 and any(not segs[0].columns[c].has_dictionary for c in q.group_by)
        oracle_threads = max(1, min(len(segs), (os.cpu_count() or 8) // max(world, 1)))
        r = native.execute(group, q, flags, prepared)
        t = r.tables[0]
        parity = "ok"

        def gather(obj):
            if world == 1:
                return [obj]
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out

        def merge_rows(parts, idx):
            exp = {}
            for part in parts:
                for k, row in part.items():
                    cur = exp.get(k)
                    if cur is None:
                        exp[k] = list(row)
                        continue
                    for i, a in enumerate(idx):
                        op = q.aggregations[a].op
                        if op in (AggOp.COUNT, AggOp.SUM):
                            cur[i] = cur[i] + row[i]
                        elif op == AggOp.MIN:
                            cur[i] = min(cur[i], row[i])
                        elif op == AggOp.MAX:
                            cur[i] = max(cur[i], row[i])
                        elif op == AggOp.AVG:
                            cur[i] = (cur[i][0] + row[i][0], cur[i][1] + row[i][1])
                        else:
                            cur[i] = set(cur[i]) | set(row[i])
            return exp

        def compare(got, exp, idx):
            assert set(got) == set(exp), f"group sets differ: {len(got)} vs {len(exp)}"
            for k, erow in exp.items():
                for i, a in enumerate(idx):
                    op, g_, e_ = q.aggregations[a].op, got[k][a], erow[i]
                    if op == AggOp.DISTINCTCOUNT:
                        e_ = len(e_)
                    if op in (AggOp.SUM, AggOp.AVG) and not cfg["exact"]:
                        gv, ev = (g_[0], e_[0]) if op == AggOp.AVG else (g_, e_)
                        assert abs(gv - ev) <= 1e-6 * max(abs(gv), abs(ev)), (k, a, g_, e_)
                    else:
                        assert g_ == e_, (k, a, g_, e_)

        try:
            many_groups = any(not segs[0].columns[c].has_dictionary for c in q.group_by)      # raw key: millions of groups
            if not many_groups:
                orc = oracle.execute_batch(oracle.PreparedBatch(segs, q), oracle_threads)
                docs = sum(g for g in gather(sum(o.stats["num_docs_scanned"] for o in orc)))
                idx = list(range(len(q.aggregations)))
                if has_dc and rows_total > 64_000_000:
                    # the value sets of 10k groups x 100k values cannot be exchanged through the host: compare everything else
                    idx = [a for a, agg in enumerate(q.aggregations) if agg.op != AggOp.DISTINCTCOUNT]
                    parity = "ok (sums / counts / statistics of every group; DISTINCTCOUNT value sets too large to exchange at this size: compared at --scale 0.02)"
                local = oracle.combine(orc)             # key -> row (DISTINCTCOUNT: value sets)
                exp = merge_rows(gather({k: [row[a] for a in idx] for k, row in local.items()}), idx)
                compare(t.rows(), exp, idx)
                del orc
            else:
                # millions of groups: (1) checksums over ALL groups -- sum of COUNT(*) = rows, sum of SUM(m) = the keyless
                # oracle's total -- and (2) 512 groups of this rank's result compared exactly with the oracle's
                keyless = parse_sql("SELECT " + ", ".join(f"{a.op.name}({a.column or '*'})" for a in q.aggregations) + " FROM t")
                ko = oracle.execute_batch(oracle.PreparedBatch(segs, keyless), oracle_threads)
                tot = [0.0] * len(q.aggregations)
                for o in ko:
                    for a, agg in enumerate(q.aggregations):
                        tot[a] += float(o.longs[a][0]) if agg.op == AggOp.COUNT else float(o.doubles[a][0])
                tot = np.sum(np.array(gather(tot)), axis=0)
                docs = int(sum(gather(sum(o.stats["num_docs_scanned"] for o in ko))))
                mine_tot = [float(np.sum(t.longs[a], dtype=np.float64)) if agg.op == AggOp.COUNT else float(np.sum(t.doubles[a], dtype=np.float64))
                            for a, agg in enumerate(q.aggregations)]
                got_tot = np.sum(np.array(gather(mine_tot)), axis=0) if hash_partitioned else np.array(mine_tot)
                for a, agg in enumerate(q.aggregations):
                    assert agg.op in (AggOp.COUNT, AggOp.SUM), "checksum parity covers COUNT / SUM"
                    assert got_tot[a] == tot[a], (agg, got_tot[a], tot[a])
                keys = t.key_values[0]
                pick = keys[:: max(1, len(keys) // 512)][:512]
                col = q.group_by[0]
                # (the sampled keys differ per rank: run the oracle for each rank's sample in turn)
                picks = gather([int(v) for v in pick])
                got_rows = t.rows()
                for src in range(world):
                    if not picks[src]:
                        continue
                    sq_src = parse_sql("SET numGroupsLimit = 20000000; SELECT " + col + ", " + ", ".join(f"{a.op.name}({a.column or '*'})" for a in q.aggregations)
                                       + f" FROM t WHERE {col} IN ({', '.join(str(v) for v in picks[src])}) GROUP BY {col} LIMIT 1000000")
                    so = oracle.execute_batch(oracle.PreparedBatch(segs, sq_src), oracle_threads)
                    idx = list(range(len(q.aggregations)))
                    exp = merge_rows(gather({k: list(row) for k, row in oracle.combine(so).items()}), idx)
                    if src == rank:
                        assert len(exp) == len(picks[src])
                        for k, erow in exp.items():
                            assert got_rows[k] == erow, (k, got_rows[k], erow)
                parity = f"ok (checksums over all {'partitions' if hash_partitioned else 'groups'}: COUNT / SUM totals exact; 512 groups per rank compared exactly with the oracle)"
            assert t.stats["num_docs_scanned"] == docs, (t.stats["num_docs_scanned"], docs)
            assert t.stats["num_total_docs"] == rows_total
        except AssertionError as e:
            parity = "MISMATCH: " + str(e)[:300]
        r.free()

        for _ in range(3):
            native.execute(group, q, flags, prepared).free()
        walls, kern, filt, agg, comm, host = [], [], [], [], [], []
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_all = time.perf_counter()
        for _ in range(args.steps):
            ts = time.perf_counter()
            r = native.execute(group, q, flags, prepared)
            walls.append(time.perf_counter() - ts)
            kern.append(r.scan_ms())
            f_, a_ = r.phase_ms()
            filt.append(f_); agg.append(a_); comm.append(r.comm_ms()); host.append(r.host_timing_us())
            ng = r.tables[0].num_groups
            matched = r.tables[0].stats["num_docs_scanned"]
            launches = r.kernel_launches
            r.free()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        elapsed = time.perf_counter() - t_all
        if world > 1:
            tt = torch.tensor([elapsed, float(np.mean(kern)), float(np.mean(comm))], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            elapsed, k_ms, c_ms = (float(x) for x in tt.tolist())
            ngt = torch.tensor([float(ng)], dtype=torch.float64, device="cuda")
            dist.all_reduce(ngt, op=dist.ReduceOp.SUM if hash_partitioned else dist.ReduceOp.MAX)
            ng = int(ngt.item())
            pv = [None] * world
            dist.all_gather_object(pv, parity)
            bad = [x for x in pv if not x.startswith("ok")]
            parity = bad[0] if bad else parity
        else:
            k_ms, c_ms = float(np.mean(kern)), 0.0
        wall = elapsed / args.steps
        if rank == 0:
            line = {"config": cfg["id"], "workload": cfg["name"], "n_gpus": world, "gpus_of_config": cfg["gpus"], "rows": rows_total,
                    "segments_per_gpu": len(segs), "groups": int(ng), "docs_matched": int(matched),
                    "rows_per_s_wall": rows_total / wall, "wall_ms_per_step": 1000 * wall, "wall_ms_median_rank0": 1000 * float(np.median(walls)),
                    "kernel_ms": k_ms, "filter_kernel_ms": float(np.mean(filt)), "agg_kernel_ms": float(np.mean(agg)), "nccl_merge_ms": c_ms,
                    "host_us_by_phase": [round(float(x), 1) for x in np.mean(np.array(host), axis=0)],
                    "rows_per_s_kernels": rows_total / (k_ms * 1e-3) if k_ms else None, "launches": int(launches),
                    "algorithmic_bytes": alg, "frac_of_measured_hbm_kernels": alg / (k_ms * 1e-3) / 1e9 / (peak * world) if k_ms else None,
                    "frac_of_measured_hbm_wall": alg / wall / 1e9 / (peak * world),
                    "parity_vs_oracle": parity, "gen_s": round(gen_s, 1), "sql": cfg["sql"](segs)[:200]}
            print(json.dumps(line), flush=True)
        group.release()
        for s in staged:
            s.release()
        del segs, staged
    if world > 1:
        import threading
        threading.Thread(target=lambda: (time.sleep(20), os._exit(0)), daemon=True).start()    # (a teardown that cannot complete must not keep the job alive)
        native.comm_destroy()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
