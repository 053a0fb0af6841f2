"""enableNullHandling on the oracle: three-valued filters (BaseFilterOperator.getTrues / getNulls / getFalses and the And / Or /
Not / BaseColumnFilterOperator overrides) and null-skipping aggregation (NullableSingleInputAggregationFunction), pinned to the
literal fixtures of the reference's NullHandlingEnabledQueriesTest (:925-1083: seven two-column rows, row counts per filter)
and to the closed-form data of NullEnabledQueriesTest (:93-122 createRecords, :281-330 and :473-495 expectations).  CPU only."""
import numpy as np
import pytest

from oracle import oracle
from pinot_b200.query import parse_sql
from pinot_b200.segment_writer import DataType, build_column, make_segment, with_nulls

INT_NULL = np.iinfo(np.int32).min          # default null value of an INT dimension (FieldSpec.DEFAULT_DIMENSION_NULL_VALUE_OF_INT)
NH = "SET enableNullHandling=true; "


def two_column_fixture(sorted_c1=False):
    """insertRowWithTwoColumns(...) x 7 of NullHandlingEnabledQueriesTest.testOrFiltering / testNotAndFiltering / testNotOrFiltering"""
    rows = [(None, None), (None, 1), (1, -1), (-1, None), (-1, 1), (1, None), (None, -1)]
    c1 = np.array([INT_NULL if a is None else a for a, _ in rows], dtype=np.int32)
    c2 = np.array([INT_NULL if b is None else b for _, b in rows], dtype=np.int32)
    n1 = np.array([a is None for a, _ in rows]); n2 = np.array([b is None for _, b in rows])
    seg = make_segment("nh7", [with_nulls(build_column("c1", DataType.INT, c1), n1), with_nulls(build_column("c2", DataType.INT, c2), n2)])
    return seg, rows


def one_column_fixture(values, sort=False):
    v = np.array([INT_NULL if x is None else x for x in values], dtype=np.int32)
    nulls = np.array([x is None for x in values])
    if sort:                                   # setSortedColumn: the segment is built in value order (nulls carry the default null value)
        order = np.argsort(v, kind="stable")
        v, nulls = v[order], nulls[order]
    return make_segment("nh1", [with_nulls(build_column("c1", DataType.INT, v), nulls)])


def count(seg, where, nh=True):
    q = parse_sql((NH if nh else "") + f"SELECT COUNT(*) FROM t WHERE {where}")
    return int(oracle.execute(seg, q).longs[0][0])


def test_reference_filter_fixtures():
    seg, rows = two_column_fixture()
    assert count(seg, "c1 > 0 OR c2 < 0") == 3                      # testOrFiltering :987-1010
    assert count(seg, "NOT (c1 > 0 AND c2 < 0)") == 3               # testNotAndFiltering :1033-1056
    assert count(seg, "NOT (c1 > 0 OR c2 < 0)") == 1                # testNotOrFiltering :1059-1083: only (-1, 1)
    assert count(one_column_fixture([None, -1, 1]), "NOT (c1 = 1)") == 1                       # testNotFiltering :1013-1030: only -1
    assert count(one_column_fixture([-1, None], sort=True), "c1 < 0") == 1                     # testRangeFiltering :947-964 (sorted column)
    assert count(one_column_fixture([None, INT_NULL], sort=True), f"c1 = {INT_NULL}") == 1     # testEqualFiltering :967-984
    # without the option the default null value is just a value (two-valued logic)
    assert count(seg, "NOT (c1 > 0 OR c2 < 0)", nh=False) == 2          # Integer.MIN_VALUE < 0 is simply true
    assert count(one_column_fixture([-1, None], sort=True), "c1 < 0", nh=False) == 2


def test_three_valued_logic_against_a_python_model():
    """Kleene logic evaluated row by row (None = unknown) against the operator-tree construction, on random trees.  Only
    trees whose And / Or children are leaves or NOTs of leaves are generated with nullable leaves below a NOT: the reference
    takes nulls_i from DIRECT column-leaf children only (And / Or / Not do not override getNulls()), so deeper shapes follow
    the operators, not Kleene -- those are covered by the device-vs-oracle fuzz."""
    rng = np.random.default_rng(2)
    n = 400
    a = rng.integers(-3, 4, n); b = rng.integers(-3, 4, n)
    an = rng.random(n) < 0.3; bn = rng.random(n) < 0.3
    seg = make_segment("k", [with_nulls(build_column("a", DataType.INT, np.where(an, INT_NULL, a).astype(np.int32)), an),
                             with_nulls(build_column("b", DataType.INT, np.where(bn, INT_NULL, b).astype(np.int32)), bn)])
    A = [None if an[i] else int(a[i]) for i in range(n)]; B = [None if bn[i] else int(b[i]) for i in range(n)]
    def k_not(x): return None if x is None else (not x)
    def k_and(x, y): return False if (x is False or y is False) else (None if (x is None or y is None) else True)
    def k_or(x, y): return True if (x is True or y is True) else (None if (x is None or y is None) else False)
    leaf = lambda col, f: [None if v is None else f(v) for v in col]
    cases = {
        "a > 0": leaf(A, lambda v: v > 0),
        "NOT (a > 0)": [k_not(x) for x in leaf(A, lambda v: v > 0)],
        "a > 0 AND b < 1": [k_and(x, y) for x, y in zip(leaf(A, lambda v: v > 0), leaf(B, lambda v: v < 1))],
        "NOT (a > 0 AND b < 1)": [k_not(k_and(x, y)) for x, y in zip(leaf(A, lambda v: v > 0), leaf(B, lambda v: v < 1))],
        "NOT (a IN (1, 2) OR b = 0)": [k_not(k_or(x, y)) for x, y in zip(leaf(A, lambda v: v in (1, 2)), leaf(B, lambda v: v == 0))],
        "a IS NULL OR b > 0": [k_or(x is None, y) for x, y in zip(A, leaf(B, lambda v: v > 0))],
        "NOT (a IS NULL) AND NOT (b <> 2)": [k_and(x is not None, k_not(y)) for x, y in zip(A, leaf(B, lambda v: v != 2))],
        "a NOT IN (0, 1) AND b BETWEEN -1 AND 1": [k_and(x, y) for x, y in zip(leaf(A, lambda v: v not in (0, 1)), leaf(B, lambda v: -1 <= v <= 1))],
    }
    for where, truth in cases.items():
        assert count(seg, where) == sum(1 for t in truth if t is True), where
    # two-valued regression: an AND over a NOT child must not be asked for more docs after EOF (DocIdSetOperator.java:63)
    stored_a, stored_b = np.where(an, INT_NULL, a), np.where(bn, INT_NULL, b)
    assert count(seg, "a > -5 AND NOT (b <> 2)", nh=False) == int(((stored_a > -5) & (stored_b == 2)).sum())


def records_fixture(base, dtype, generate_nulls=True, num_records=1000):
    """NullEnabledQueriesTest.createRecords (:93-122): value = base + i for even i (key 1 below NUM_RECORDS / 2, else 2), null
    for odd i.  The key column is NOT nullable here (null keys are outside the offloaded set): null rows carry key 0."""
    vals, keys, nulls = [], [], []
    for i in range(num_records):
        if i % 2 == 0:
            vals.append(base + i); keys.append(1 if i < num_records // 2 else 2); nulls.append(False)
        elif generate_nulls:
            vals.append(0); keys.append(0); nulls.append(True)      # default null value of a metric: 0
    np_t = {DataType.INT: np.int32, DataType.LONG: np.int64, DataType.FLOAT: np.float32, DataType.DOUBLE: np.float64}[dtype]
    v = np.array(vals).astype(np_t)
    col = with_nulls(build_column("column", dtype, v, dictionary=False), np.array(nulls))
    return make_segment("rec", [col, build_column("key", DataType.INT, np.array(keys, dtype=np.int32))]), v, np.array(keys), np.array(nulls)


@pytest.mark.parametrize("dtype,base", [(DataType.INT, 7), (DataType.LONG, 1 << 40), (DataType.FLOAT, 0.25), (DataType.DOUBLE, 0.6180339887)])
@pytest.mark.parametrize("generate_nulls", [True, False])
def test_reference_aggregation_expectations(dtype, base, generate_nulls):
    seg, v, keys, nulls = records_fixture(base, dtype, generate_nulls)
    nn = ~nulls
    total = float(np.sum(v[nn].astype(np.float64)))
    # :473-495  SELECT COUNT(col), MIN(col), MAX(col), AVG(col), SUM(col)  (per segment: count 500, min base, max base + 998)
    r = oracle.execute(seg, parse_sql(NH + "SELECT COUNT(column), MIN(column), MAX(column), AVG(column), SUM(column) FROM t"))
    assert int(r.longs[0][0]) == 500
    assert abs(r.doubles[1][0] - float(v[0])) < 1e-1 and abs(r.doubles[2][0] - float(v[nn][-1])) < 1e-1
    assert int(r.longs[3][0]) == 500 and abs(r.doubles[3][0] / r.longs[3][0] - total / 500) < 1e-1
    assert abs(r.doubles[4][0] - total) < 1e-1 * max(1.0, abs(total) * 1e-12) and int(r.longs[4][0]) == 500
    # :281-330  SUM / MIN / MAX / COUNT(col) GROUP BY key: keys 1 and 2 hold 250 values each; the rows whose value is null form a
    # group of their own whose SUM / MIN / MAX are NULL and whose COUNT(col) is 0 ("similar to Presto")
    g = oracle.execute(seg, parse_sql(NH + "SELECT key, SUM(column), MIN(column), MAX(column), COUNT(column) FROM t GROUP BY key LIMIT 10"))
    rows = {k[0]: i for i, k in enumerate(g.decoded_keys())}
    assert sorted(rows) == ([0, 1, 2] if generate_nulls else [1, 2])
    for key in (1, 2):
        i = rows[key]; sel = nn & (keys == key)
        assert int(g.longs[3][i]) == 250 and int(g.longs[0][i]) == 250
        assert abs(g.doubles[0][i] - float(np.sum(v[sel].astype(np.float64)))) < 1e-1 * max(1.0, abs(total) * 1e-12)
        assert abs(g.doubles[1][i] - float(v[sel][0])) < 1e-1 and abs(g.doubles[2][i] - float(v[sel][-1])) < 1e-1
    if generate_nulls:
        i = rows[0]
        assert [int(g.longs[a][i]) for a in range(4)] == [0, 0, 0, 0]          # no input seen: SUM / MIN / MAX are SQL NULL, COUNT(col) = 0
    # the same query without the option: the default null value 0 takes part
    r0 = oracle.execute(seg, parse_sql("SELECT COUNT(column), MIN(column), SUM(column) FROM t"))
    assert int(r0.longs[0][0]) == len(v) and (not generate_nulls or r0.doubles[1][0] == 0.0) and abs(r0.doubles[2][0] - total) < 1e-1 * max(1.0, abs(total) * 1e-12)


def test_all_null_column_returns_null():
    """AllNullQueriesTest :443-468: COUNT(col) = 0 and MIN / MAX / AVG / SUM are NULL when every value is null."""
    n = 300
    seg = make_segment("alln", [with_nulls(build_column("column", DataType.LONG, np.zeros(n, dtype=np.int64)), np.ones(n, dtype=bool)),
                                build_column("d", DataType.INT, np.arange(n, dtype=np.int32) % 3)])
    r = oracle.execute(seg, parse_sql(NH + "SELECT COUNT(column), MIN(column), MAX(column), AVG(column), SUM(column) FROM t"))
    assert [int(r.longs[a][0]) for a in range(5)] == [0, 0, 0, 0, 0]
    from tests.parity import oracle_rows
    assert oracle_rows(r)[()] == [0, None, None, None, None]
    g = oracle.execute(seg, parse_sql(NH + "SELECT d, SUM(column), DISTINCTCOUNT(column) FROM t GROUP BY d LIMIT 10"))
    assert g.num_groups == 3 and all(int(x) == 0 for x in g.longs[0]) and all(int(x) == 0 for x in g.longs[1])


def test_combine_of_null_results():
    from tests.parity import combined_rows
    rng = np.random.default_rng(6)
    segs = []
    for s in range(3):
        n = 500
        d = rng.integers(0, 4, n).astype(np.int32)
        x = rng.integers(1, 100, n).astype(np.int64)
        nulls = (d == s) | (rng.random(n) < 0.2)              # group s is entirely null in segment s
        segs.append(make_segment(f"s{s}", [build_column("d", DataType.INT, d), with_nulls(build_column("x", DataType.LONG, np.where(nulls, 0, x)), nulls)]))
    q = parse_sql(NH + "SELECT d, SUM(x), MIN(x), AVG(x), COUNT(x), COUNT(*) FROM t GROUP BY d LIMIT 10")
    table = oracle.combine([oracle.execute(s, q) for s in segs])
    for key, row in table.items():
        assert row[0] is not None and row[1] is not None and row[2][1] == row[3] and row[4] >= row[3]     # every group has non-null inputs somewhere


def test_null_skipping_aggregations_against_pandas():
    """SUM / MIN / MAX / AVG / COUNT(col) / COUNT(*) / DISTINCTCOUNT with nullable inputs against pandas (skipna semantics,
    min_count=1 for SUM so that an all-null group is NaN = SQL NULL), group-by on a non-nullable key, filters whose leaves sit
    directly under AND / OR / NOT (Kleene logic = the operators' logic there)."""
    pd = pytest.importorskip("pandas")
    rng = np.random.default_rng(11)
    n = 6000
    g = rng.integers(0, 9, n).astype(np.int32)
    x = rng.integers(-50, 50, n).astype(np.int64); xn = rng.random(n) < 0.35
    y = np.round(rng.normal(0, 4, n), 1); yn = rng.random(n) < 0.2
    xn[g == 4] = True                                            # group 4: every x is null
    seg = make_segment("pd", [build_column("g", DataType.INT, g),
                              with_nulls(build_column("x", DataType.LONG, np.where(xn, 0, x), dictionary=False), xn),
                              with_nulls(build_column("y", DataType.DOUBLE, np.where(yn, 0.0, y)), yn)])
    df = pd.DataFrame({"g": g, "x": np.where(xn, np.nan, x.astype(np.float64)), "y": np.where(yn, np.nan, y)})
    filters = {
        None: np.ones(n, dtype=bool),
        "y > -2": (df.y > -2).fillna(False).to_numpy(),
        "NOT (y > -2)": ((df.y <= -2)).fillna(False).to_numpy(),                                     # null y: unknown, dropped
        "x < 10 OR y > 3": ((df.x < 10).fillna(False) | (df.y > 3).fillna(False)).to_numpy(),
        "NOT (x < 10 AND y > 0)": ((df.x >= 10).fillna(False) | (df.y <= 0).fillna(False)).to_numpy(),   # false iff some operand is false
        "x IS NULL AND NOT (y IS NULL)": (df.x.isna() & df.y.notna()).to_numpy(),
    }
    for where, mask in filters.items():
        sql = NH + "SELECT g, COUNT(*), COUNT(x), SUM(x), MIN(x), MAX(y), AVG(y), DISTINCTCOUNT(x) FROM t" + (f" WHERE {where}" if where else "") + " GROUP BY g LIMIT 100"
        r = oracle.execute(seg, parse_sql(sql))
        sub = df[mask]
        exp = sub.groupby("g").agg(cnt=("g", "size"), cx=("x", "count"), sx=("x", lambda s: s.sum(min_count=1)), mn=("x", "min"), mx=("y", "max"),
                                   avg=("y", "mean"), cy=("y", "count"), dc=("x", "nunique"))
        keys = [k[0] for k in r.decoded_keys()]
        assert sorted(keys) == sorted(exp.index.tolist()), where
        for i, k in enumerate(keys):
            e = exp.loc[k]
            assert int(r.longs[0][i]) == int(e.cnt) and int(r.longs[1][i]) == int(e.cx), (where, k)
            assert int(r.longs[2][i]) == int(e.cx) and (np.isnan(e.sx) if e.cx == 0 else r.doubles[2][i] == e.sx), (where, k)
            assert int(r.longs[3][i]) == int(e.cx) and (e.cx == 0 or r.doubles[3][i] == e.mn), (where, k)
            assert int(r.longs[4][i]) == int(e.cy) and (e.cy == 0 or r.doubles[4][i] == e.mx), (where, k)
            assert int(r.longs[5][i]) == int(e.cy) and (e.cy == 0 or abs(r.doubles[5][i] / r.longs[5][i] - e.avg) < 1e-9), (where, k)
            assert int(r.longs[6][i]) == int(e.dc), (where, k)
    # keyless, everything filtered out: COUNT 0, every other function NULL
    r = oracle.execute(seg, parse_sql(NH + "SELECT COUNT(*), COUNT(x), SUM(x), MIN(y) FROM t WHERE y > 1000"))
    assert [int(r.longs[a][0]) for a in range(4)] == [0, 0, 0, 0]
