"""The oracle (the checker of every GPU parity test) against an independent pandas restatement of group-by / aggregation on
random queries — on top of the reference's golden literals in test_oracle_golden.py.  CPU only."""
import numpy as np
import pandas as pd
import pytest

from oracle import oracle
from pinot_b200 import datagen
from pinot_b200.query import AggOp, parse_sql
from tests.test_cpu_lowering_fuzz import _column_values, _expr, evaluate_sql

COLS = ["c1", "c3", "d0", "d1", "d3", "s0", "t0", "m0", "m1", "x0", "k0"]


@pytest.fixture(scope="module")
def table():
    seg = datagen.make_segment_synth(11, 120_011, columns=COLS)
    frame = pd.DataFrame({c: ([bytes(x) for x in v] if v.dtype.kind == "S" else v) for c, v in ((c, _column_values(seg, c)) for c in COLS)})
    return seg, frame


@pytest.mark.parametrize("seed", range(60))
def test_group_by_aggregations_vs_pandas(table, seed):
    seg, frame = table
    rng = np.random.default_rng(500 + seed)
    keys = list(rng.choice(["d0", "d1", "d3", "s0", "t0", "k0", "c3"], size=int(rng.integers(0, 4)), replace=False))
    metrics = ["m0", "m1", "x0", "k0", "c1"]
    aggs = []
    for _ in range(int(rng.integers(1, 5))):
        op = str(rng.choice(["COUNT", "SUM", "MIN", "MAX", "AVG", "DISTINCTCOUNT"]))
        col = str(rng.choice(["c1", "d3", "c3"])) if op == "DISTINCTCOUNT" else str(rng.choice(metrics))
        aggs.append((op, None if op == "COUNT" else col))
    where = _expr(rng, seg, ["c1", "c3", "d0", "t0", "x0", "s0"], depth=1) if rng.random() < 0.8 else None
    select = ", ".join(f"{op}({col or '*'})" for op, col in aggs)
    sql = f"SET numGroupsLimit = 10000000; SELECT {select} FROM t" + (f" WHERE {where}" if where else "") + \
          (f" GROUP BY {', '.join(keys)} LIMIT 10000000" if keys else "")
    q = parse_sql(sql)
    r = oracle.execute(seg, q)
    sub = frame[evaluate_sql(seg, q.filter)] if q.filter is not None else frame
    assert r.stats["num_docs_scanned"] == len(sub), sql
    got_keys = [tuple(x) for x in r.decoded_keys()] if keys else [()]
    grouped = sub.groupby(keys, sort=False) if keys else None
    if keys:
        sizes = grouped.size()
        exp_keys = [k if isinstance(k, tuple) else (k,) for k in sizes.index.tolist()]
        assert sorted(map(repr, got_keys)) == sorted(map(repr, exp_keys)), sql
        pos = {k: i for i, k in enumerate(exp_keys)}
        order = np.array([pos[k] for k in got_keys], dtype=np.int64)          # oracle row -> pandas row
    for a, (op, col) in enumerate(aggs):
        if keys:
            if op == "COUNT":
                exp = sizes.to_numpy()
            elif op == "DISTINCTCOUNT":
                exp = grouped[col].nunique().to_numpy()
            else:
                exp = getattr(grouped[col], {"SUM": "sum", "AVG": "sum", "MIN": "min", "MAX": "max"}[op])().to_numpy(dtype=np.float64)
            exp = exp[order]
            cnt = sizes.to_numpy()[order]
        else:
            n = len(sub)
            cnt = np.array([n])
            if op == "COUNT":
                exp = np.array([n])
            elif op == "DISTINCTCOUNT":
                exp = np.array([sub[col].nunique()])
            elif n == 0:                                                 # keyless query without matching docs: the defaults
                exp = np.array([{"SUM": 0.0, "AVG": 0.0, "MIN": np.inf, "MAX": -np.inf}[op]])
            else:
                v = sub[col].to_numpy(dtype=np.float64)
                exp = np.array([{"SUM": v.sum(), "AVG": v.sum(), "MIN": v.min(), "MAX": v.max()}[op]])
        if op in ("COUNT", "DISTINCTCOUNT"):
            assert (r.longs[a] == exp).all(), (sql, op, col)
        else:
            assert np.allclose(r.doubles[a], exp, rtol=1e-9, atol=0.0, equal_nan=False), (sql, op, col)
            if op == "AVG":
                assert (r.longs[a] == cnt).all(), (sql, op, col)


@pytest.mark.parametrize("seed", range(30))
def test_filtered_aggregations_vs_pandas(table, seed):
    """Random FILTER(WHERE ...) clauses: every group of the main filter exists; a function sees only the docs that pass its
    clause and keeps its default where none does (FilteredGroupByOperator.java:108-159)."""
    seg, frame = table
    rng = np.random.default_rng(9000 + seed)
    keys = list(rng.choice(["d0", "d1", "d3", "t0"], size=int(rng.integers(0, 3)), replace=False))
    clause_cols = ["c1", "c3", "d0", "x0", "t0"]
    clauses = [_expr(rng, seg, clause_cols, depth=1) for _ in range(int(rng.integers(1, 4)))]
    aggs = []
    for _ in range(int(rng.integers(1, 5))):
        op = str(rng.choice(["COUNT", "SUM", "MIN", "MAX", "AVG", "DISTINCTCOUNT"]))
        col = str(rng.choice(["c1", "d3"])) if op == "DISTINCTCOUNT" else str(rng.choice(["m0", "m1", "x0", "k0"]))
        clause = int(rng.integers(-1, len(clauses)))
        aggs.append((op, None if op == "COUNT" else col, clause))
    where = _expr(rng, seg, ["c1", "d0", "x0"], depth=1) if rng.random() < 0.7 else None
    select = ", ".join(f"{op}({col or '*'})" + (f" FILTER(WHERE {clauses[cl]})" if cl >= 0 else "") for op, col, cl in aggs)
    sql = f"SELECT {select} FROM t" + (f" WHERE {where}" if where else "") + (f" GROUP BY {', '.join(keys)} LIMIT 10000000" if keys else "")
    q = parse_sql(sql)
    r = oracle.execute(seg, q)
    main = evaluate_sql(seg, q.filter) if q.filter is not None else np.ones(seg.num_docs, bool)
    sub = frame[main]
    got_keys = [tuple(x) for x in r.decoded_keys()] if keys else [()]
    if keys:
        exp_keys = [k if isinstance(k, tuple) else (k,) for k in sub.groupby(keys, sort=False).size().index.tolist()]
        assert sorted(map(repr, got_keys)) == sorted(map(repr, exp_keys)), sql
    defaults = {"COUNT": 0, "DISTINCTCOUNT": 0, "SUM": 0.0, "AVG": 0.0, "MIN": np.inf, "MAX": -np.inf}
    for a, (op, col, cl) in enumerate(aggs):
        lane = frame[main & evaluate_sql(seg, q.aggregations[a].filter)] if cl >= 0 else sub
        if keys:
            g = lane.groupby(keys, sort=False)
            if op == "COUNT":
                ser = g.size()
            elif op == "DISTINCTCOUNT":
                ser = g[col].nunique()
            else:
                ser = getattr(g[col], {"SUM": "sum", "AVG": "sum", "MIN": "min", "MAX": "max"}[op])()
            lut = {(k if isinstance(k, tuple) else (k,)): v for k, v in ser.items()}
            cnt = {(k if isinstance(k, tuple) else (k,)): v for k, v in g.size().items()}
        else:
            n = len(lane)
            v = lane[col].to_numpy(dtype=np.float64) if col and op not in ("COUNT", "DISTINCTCOUNT") and n else None
            val = n if op == "COUNT" else (lane[col].nunique() if op == "DISTINCTCOUNT" else
                                           (defaults[op] if n == 0 else {"SUM": v.sum(), "AVG": v.sum(), "MIN": v.min(), "MAX": v.max()}[op]))
            lut, cnt = {(): val}, {(): n}
        for gi, key in enumerate(got_keys):
            exp = lut.get(key, defaults[op])
            if op in ("COUNT", "DISTINCTCOUNT"):
                assert r.longs[a][gi] == exp, (sql, key, op)
            else:
                assert r.doubles[a][gi] == pytest.approx(exp, rel=1e-9), (sql, key, op, col)
                if op == "AVG":
                    assert r.longs[a][gi] == cnt.get(key, 0), (sql, key)


@pytest.mark.parametrize("seed", range(12))
def test_cross_segment_merge_vs_pandas(seed):
    """oracle.combine (GroupByCombineOperator / IndexedTable merge by decoded key: the checker of PB_Q_COMBINE) against pandas
    over the concatenated rows of three segments whose dimension dictionaries differ."""
    rng = np.random.default_rng(40 + seed)
    cols = ["c1", "d1", "d3", "s0", "m0", "x0"]
    segs = [datagen.make_segment_synth(20 + i, n, columns=cols, vary_dim_dictionaries=True) for i, n in enumerate((30_011, 20_003, 12_345))]
    frames = [pd.DataFrame({c: ([bytes(x) for x in v] if v.dtype.kind == "S" else v) for c, v in ((c, _column_values(s, c)) for c in cols)})
              for s in segs]
    keys = list(rng.choice(["d1", "d3", "s0"], size=int(rng.integers(1, 3)), replace=False))
    where = _expr(rng, segs[0], ["c1", "x0"], depth=1)
    q = parse_sql(f"SET numGroupsLimit = 10000000; SELECT COUNT(*), SUM(m0), MIN(x0), MAX(m0), AVG(x0), DISTINCTCOUNT(c1) FROM t "
                  f"WHERE {where} GROUP BY {', '.join(keys)} LIMIT 10000000")
    merged = oracle.combine([oracle.execute(s, q) for s in segs])
    sub = pd.concat([f[evaluate_sql(s, q.filter)] for s, f in zip(segs, frames)], ignore_index=True)
    g = sub.groupby(keys, sort=False)
    exp = pd.DataFrame({"n": g.size(), "sum": g["m0"].sum(), "min": g["x0"].min(), "max": g["m0"].max(), "avg": g["x0"].sum(), "dc": g["c1"].nunique()})
    assert len(merged) == len(exp)
    for k, row in exp.iterrows():
        key = k if isinstance(k, tuple) else (k,)
        got = merged[key]
        assert got[0] == row["n"] and got[1] == float(row["sum"]) and got[2] == row["min"] and got[3] == float(row["max"])
        assert got[4][1] == row["n"] and got[4][0] == pytest.approx(row["avg"], rel=1e-12) and len(got[5]) == row["dc"]


def test_distinctcount_on_raw_columns_against_pandas():
    """BaseDistinctAggregateAggregationFunction.java:157-226: a raw column's DISTINCTCOUNT keeps per-group value sets"""
    import pandas as pd
    from pinot_b200.segment_writer import DataType, build_column, make_segment
    rng = np.random.Generator(np.random.PCG64(3))
    n = 20_000
    g = rng.integers(0, 12, size=n)
    ri = rng.integers(-50, 50, size=n)
    rl = rng.integers(-10**12, 10**12, size=n) // 10**11 * 10**11
    rd = np.round(rng.normal(size=n), 1)
    rd[::97] = -0.0
    seg = make_segment("rawdc", [build_column("g", DataType.INT, g), build_column("ri", DataType.INT, ri, dictionary=False),
                                 build_column("rl", DataType.LONG, rl, dictionary=False), build_column("rd", DataType.DOUBLE, rd, dictionary=False),
                                 build_column("f", DataType.INT, rng.integers(0, 100, size=n))])
    df = pd.DataFrame({"g": g, "ri": ri, "rl": rl, "rd": rd, "f": seg.columns["f"].dictionary_values()[0] * 0 + rng.integers(0, 1, size=n)})
    q = parse_sql("SELECT g, DISTINCTCOUNT(ri), DISTINCTCOUNT(rl), DISTINCTCOUNT(rd), COUNT(*) FROM t GROUP BY g LIMIT 1000")
    o = oracle.execute(seg, q)
    keys = [k[0] for k in o.decoded_keys()]
    for gi, k in enumerate(keys):
        sub = df[df.g == k]
        assert o.longs[0][gi] == sub.ri.nunique() and o.longs[1][gi] == sub.rl.nunique()
        # -0.0 and 0.0 are different members of a DoubleOpenHashSet (Double.doubleToLongBits)
        assert o.longs[2][gi] == len(set(sub.rd.values.view(np.int64).tolist()))
        off, vals = o.distinct[0]
        assert sorted(set(sub.ri.tolist())) == vals[off[gi]:off[gi + 1]].tolist()
    ok = oracle.execute(seg, parse_sql("SELECT DISTINCTCOUNT(ri), DISTINCTCOUNT(rl) FROM t WHERE g < 6"))
    assert ok.longs[0][0] == df[df.g < 6].ri.nunique() and ok.longs[1][0] == df[df.g < 6].rl.nunique()


def test_native_combine_matches_the_python_merge():
    """orc_execute_combined (worker threads fold their segments' results into IndexedTables keyed by the decoded group key:
    GroupByCombineOperator.java:132-147, IndexedTable.java:99-125) against oracle.combine over the same per-segment results,
    with segment-local dictionaries that differ, on 1 and 5 threads, repeatedly (the worker pool is reused)."""
    from pinot_b200 import datagen
    segs = [datagen.make_segment_synth(700 + i, 30_000 + 1000 * i, vary_dim_dictionaries=(i % 2 == 1)) for i in range(7)]
    q = parse_sql("SELECT d1, d2, COUNT(*), SUM(m0), MIN(m1), MAX(m2), AVG(m0) FROM t WHERE c1 < 400 GROUP BY d1, d2 LIMIT 100000")
    prep = oracle.PreparedBatch(segs, q)
    exp = oracle.combine([oracle.execute(s, q) for s in segs])
    for threads in (1, 5, 5, 3):
        keys, dbl, lng = oracle.execute_combined(prep, threads)
        assert len(keys) == len(exp)
        for k, d, l in zip(keys, dbl, lng):
            row = exp[tuple(int(x) for x in k)]
            assert int(l[0]) == row[0] and d[1] == row[1] and d[2] == row[2] and d[3] == row[3] and (d[4], int(l[4])) == row[4]
    # batches on the pool still return per-segment results
    res = oracle.execute_batch(prep, 4)
    assert [r.num_groups for r in res] == [oracle.execute(s, q).num_groups for s in segs]
    # DISTINCTCOUNT is left to the Python merge
    assert oracle.execute_combined(oracle.PreparedBatch(segs, parse_sql("SELECT d1, DISTINCTCOUNT(c1) FROM t GROUP BY d1 LIMIT 10")), 2) is None


def test_nan_and_signed_zero_in_min_max():
    """MinAggregationFunction / MaxAggregationFunction: the keyless path folds with Math.min / Math.max (a NaN input makes the
    result NaN, -0.0 < 0.0; MinAggregationFunction.java:97-124), the group-by path compares with a strict "<" against the
    holder (:163-188), which never lets a NaN in."""
    from pinot_b200.segment_writer import DataType, build_column, make_segment
    x = np.array([1.5, np.nan, -2.0, 0.0, -0.0, 7.0])
    g = np.array([0, 0, 0, 1, 1, 1], dtype=np.int32)
    seg = make_segment("nan", [build_column("x", DataType.DOUBLE, x, dictionary=False), build_column("g", DataType.INT, g)])
    r = oracle.execute(seg, parse_sql("SELECT MIN(x), MAX(x) FROM t"))
    assert np.isnan(r.doubles[0][0]) and np.isnan(r.doubles[1][0])
    r = oracle.execute(seg, parse_sql("SELECT MIN(x), MAX(x) FROM t WHERE g = 1"))
    assert r.doubles[0][0] == 0.0 and np.signbit(r.doubles[0][0]) and r.doubles[1][0] == 7.0            # Math.min(0.0, -0.0) = -0.0
    r = oracle.execute(seg, parse_sql("SELECT g, MIN(x), MAX(x) FROM t GROUP BY g LIMIT 10"))
    rows = {k[0]: (r.doubles[0][i], r.doubles[1][i]) for i, k in enumerate(r.decoded_keys())}
    assert rows[0] == (-2.0, 1.5) and rows[1][1] == 7.0 and rows[1][0] == 0.0 and not np.signbit(rows[1][0])   # strict <: the first zero stays
