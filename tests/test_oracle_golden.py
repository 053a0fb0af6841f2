"""Pin the oracle to the reference's own golden values.

Every expected number below is a literal quoted from the reference's tests (CTEST =
pinot-core/src/test/java/org/apache/pinot/queries):
  InnerSegmentAggregationSingleValueQueriesTest.java:43-177   (per-segment operator results + ExecutionStatistics)
  InterSegmentAggregationSingleValueQueriesTest.java:47-258   (4 identical segments merged)
"""
import numpy as np
import pytest

from oracle import oracle
from pinot_b200.query import parse_sql
from tests.fixtures import FILTER, sv_segment

AGG = "SELECT COUNT(*), SUM(column1), MAX(column3), MIN(column6), AVG(column7) FROM testTable"


def _inner(sql):
    seg = sv_segment()
    return oracle.execute(seg, parse_sql(sql))


def _check_agg(r, g, expected):
    assert int(r.longs[0][g]) == expected[0]
    assert int(r.doubles[1][g]) == expected[1]
    assert int(r.doubles[2][g]) == expected[2]
    assert int(r.doubles[3][g]) == expected[3]
    assert int(r.doubles[4][g]) == expected[4]
    assert int(r.longs[4][g]) == expected[5]


def _stats(r):
    s = r.stats
    return (s["num_docs_scanned"], s["num_entries_scanned_in_filter"], s["num_entries_scanned_post_filter"],
            s["num_total_docs"])


def test_aggregation_only():   # InnerSegment...Test.java:43-60
    r = _inner(AGG)
    assert _stats(r) == (30000, 0, 120000, 30000)
    _check_agg(r, 0, (30000, 32317185437847, 2147419555, 1689277, 28175373944314, 30000))
    r = _inner(AGG + FILTER)
    assert _stats(r) == (6129, 63064, 24516, 30000)
    _check_agg(r, 0, (6129, 6875947596072, 999813884, 1980174, 4699510391301, 6129))


def _find(r, key):
    keys = r.decoded_keys()
    assert key in keys, f"group {key} not found"
    return keys.index(key)


@pytest.mark.parametrize("group_by,holder,post,post_f,key,exp,key_f,exp_f", [
    # :96-111 ARRAY_BASED
    (" GROUP BY column9", 1, 150000, 30645, (11270,), (1, 815409257, 1215316262, 1328642550, 788414092, 1),
     (242920,), (3, 4348938306, 407993712, 296467636, 5803888725, 3)),
    # :115-131 INT_MAP_BASED
    (" GROUP BY column9, column11, column12", 2, 210000, 42903, (1813102948, b"P", b"HEuxNvH"),
     (4, 2062187196, 1988589001, 394608493, 4782388964, 4),
     (1176631727, b"P", b"KrNxpdycSiwoRohEiTIlLqDHnx"), (1, 716185211, 489993380, 371110078, 487714191, 1)),
    # :135-152 LONG_MAP_BASED
    (" GROUP BY column1, column6, column9, column11, column12", 3, 210000, 42903,
     (484569489, 16200443, 1159557463, b"P", b"MaztCmmxxgguBUxPti"), (2, 969138978, 995355481, 16200443, 2222394270, 2),
     (1318761745, 353175528, 1172307870, b"P", b"HEuxNvH"), (2, 2637523490, 557154208, 353175528, 2427862396, 2)),
    # :156-174 ARRAY_MAP_BASED
    (" GROUP BY column1, column3, column6, column7, column9, column11, column12, column17, column18", 4, 270000, 55161,
     (1784773968, 204243323, 628170461, 1985159279, 296467636, b"P", b"HEuxNvH", 402773817, 2047180536),
     (1, 1784773968, 204243323, 628170461, 1985159279, 1),
     (1361199163, 178133991, 296467636, 788414092, 1719301234, b"P", b"MaztCmmxxgguBUxPti", 1284373442, 752388855),
     (1, 1361199163, 178133991, 296467636, 788414092, 1)),
])
def test_group_by(group_by, holder, post, post_f, key, exp, key_f, exp_f):
    r = _inner(AGG + group_by)
    assert r.stats["key_holder"] == holder
    assert _stats(r) == (30000, 0, post, 30000)
    _check_agg(r, _find(r, key), exp)
    r = _inner(AGG + FILTER + group_by)
    assert _stats(r) == (6129, 63064, post_f, 30000)
    _check_agg(r, _find(r, key_f), exp_f)


def _inter(sql):
    """BaseQueriesTest.getBrokerResponse: 2 segment objects x 2 'servers' = 4 identical segments."""
    seg = sv_segment()
    q = parse_sql(sql)
    r = oracle.execute(seg, q)
    return oracle.combine([r, r, r, r]), q


def test_inter_segment_count():   # InterSegment...Test.java:47-88
    t, _ = _inter("SELECT COUNT(*) FROM testTable")
    assert t[()][0] == 120000
    t, _ = _inter("SELECT COUNT(*) FROM testTable" + FILTER)
    assert t[()][0] == 24516
    t, _ = _inter("SELECT COUNT(*) FROM testTable GROUP BY column9")
    assert max(v[0] for v in t.values()) == 64420
    t, _ = _inter("SELECT COUNT(*) FROM testTable" + FILTER + " GROUP BY column9")
    assert max(v[0] for v in t.values()) == 17080


def test_inter_segment_min_max():   # :92-147
    t, _ = _inter("SELECT MAX(column1), MAX(column3) FROM testTable")
    assert t[()] == [2146952047.0, 2147419555.0]
    t, _ = _inter("SELECT MAX(column1), MAX(column3) FROM testTable" + FILTER)
    assert t[()] == [2146952047.0, 999813884.0]
    t, _ = _inter("SELECT MIN(column1), MIN(column3) FROM testTable")
    assert t[()] == [240528.0, 17891.0]


def test_inter_segment_distinct_count():   # :235-258
    t, _ = _inter("SELECT DISTINCTCOUNT(column1), DISTINCTCOUNT(column3) FROM testTable")
    assert [len(s) for s in t[()]] == [6582, 21910]
    t, _ = _inter("SELECT DISTINCTCOUNT(column1), DISTINCTCOUNT(column3) FROM testTable" + FILTER)
    assert [len(s) for s in t[()]] == [1872, 4556]
    t, _ = _inter("SELECT DISTINCTCOUNT(column1), DISTINCTCOUNT(column3) FROM testTable GROUP BY column9")
    assert max(len(v[0]) for v in t.values()) == 3495
    assert max(len(v[1]) for v in t.values()) == 11961
    t, _ = _inter("SELECT DISTINCTCOUNT(column1), DISTINCTCOUNT(column3) FROM testTable" + FILTER + " GROUP BY column9")
    assert max(len(v[0]) for v in t.values()) == 1272
    assert max(len(v[1]) for v in t.values()) == 3289


FILTERED = ("SELECT SUM(column6) FILTER(WHERE column6 > 5), COUNT(*) FILTER(WHERE column1 IS NOT NULL), "
            "MAX(column3) FILTER(WHERE column3 IS NOT NULL), SUM(column3), AVG(column7) FILTER(WHERE column7 > 0) FROM testTable")
FILTERED_3 = ("SELECT SUM(column6) FILTER(WHERE column6 > 5 OR column6 < 15), COUNT(*) FILTER(WHERE column1 IS NOT NULL), "
              "MAX(column3) FILTER(WHERE column3 IS NOT NULL AND column3 > 0), SUM(column3), "
              "AVG(column7) FILTER(WHERE column7 > 0 AND column7 < 100) FROM testTable")


def filtered_row(r, g=0):
    return (int(r.doubles[0][g]), int(r.longs[1][g]), int(r.doubles[2][g]), int(r.doubles[3][g]), int(r.doubles[4][g]), int(r.longs[4][g]))


def test_filtered_aggregations():   # InnerSegment...Test.java:62-93 (FilteredAggregationOperator, one swim-lane per FILTER clause)
    r = _inner(FILTERED + " WHERE column3 > 0")
    assert _stats(r) == (150000, 0, 120000, 30000)
    assert filtered_row(r) == (22266008882250, 30000, 2147419555, 32289159189150, 28175373944314, 30000)
    r = _inner(FILTERED)
    assert _stats(r) == (150000, 0, 120000, 30000)
    assert filtered_row(r) == (22266008882250, 30000, 2147419555, 32289159189150, 28175373944314, 30000)
    r = _inner(FILTERED_3)
    assert _stats(r) == (120000, 0, 90000, 30000)
    assert filtered_row(r) == (22266008882250, 30000, 2147419555, 32289159189150, 0, 0)


def test_filtered_aggregations_equal_separate_queries():
    """FilteredAggregationsTest.java's property: AGG(x) FILTER(WHERE p) under WHERE m == AGG(x) WHERE m AND p, keyless
    and grouped (every group of the main filter is present; functions without a matching doc keep their default)."""
    seg = sv_segment()
    main, p1, p2 = "column6 < 1500000000", "column1 > 100000000 AND column11 <> 'P'", "column7 IN (1111197135, 296467636, 675163196) OR column17 < 100000000"
    q = parse_sql(f"SELECT SUM(column1) FILTER(WHERE {p1}), COUNT(*) FILTER(WHERE {p2}), MIN(column3) FILTER(WHERE {p1}), COUNT(*), "
                  f"AVG(column6) FILTER(WHERE {p2}), DISTINCTCOUNT(column9) FILTER(WHERE {p1}), MAX(column18) FROM testTable WHERE {main}")
    r = oracle.execute(seg, q)
    a = oracle.execute(seg, parse_sql(f"SELECT SUM(column1), MIN(column3), DISTINCTCOUNT(column9), COUNT(*) FROM testTable WHERE {main} AND ({p1})"))
    b = oracle.execute(seg, parse_sql(f"SELECT COUNT(*), AVG(column6) FROM testTable WHERE {main} AND ({p2})"))
    c = oracle.execute(seg, parse_sql(f"SELECT COUNT(*), MAX(column18) FROM testTable WHERE {main}"))
    assert r.doubles[0][0] == a.doubles[0][0] and r.doubles[2][0] == a.doubles[1][0] and r.longs[5][0] == a.longs[2][0]
    assert r.longs[1][0] == b.longs[0][0] and r.doubles[4][0] == b.doubles[1][0] and r.longs[4][0] == b.longs[1][0]
    assert r.longs[3][0] == c.longs[0][0] and r.doubles[6][0] == c.doubles[1][0]
    assert r.stats["num_docs_scanned"] == a.longs[3][0] + b.longs[0][0] + c.longs[0][0]
    # grouped
    gq = parse_sql(f"SELECT column11, SUM(column1) FILTER(WHERE {p1}), COUNT(*) FILTER(WHERE {p2}), MAX(column3) FILTER(WHERE {p1}) "
                   f"FROM testTable WHERE {main} GROUP BY column11")
    g = oracle.execute(seg, gq)
    gm = oracle.execute(seg, parse_sql(f"SELECT column11, COUNT(*) FROM testTable WHERE {main} GROUP BY column11"))
    ga = oracle.execute(seg, parse_sql(f"SELECT column11, SUM(column1), MAX(column3) FROM testTable WHERE {main} AND ({p1}) GROUP BY column11"))
    gb = oracle.execute(seg, parse_sql(f"SELECT column11, COUNT(*) FROM testTable WHERE {main} AND ({p2}) GROUP BY column11"))
    assert sorted(g.decoded_keys()) == sorted(gm.decoded_keys())          # the main lane creates every group
    ka = {k: i for i, k in enumerate(ga.decoded_keys())}
    kb = {k: i for i, k in enumerate(gb.decoded_keys())}
    for i, k in enumerate(g.decoded_keys()):
        assert g.doubles[0][i] == (ga.doubles[0][ka[k]] if k in ka else 0.0)
        assert g.doubles[2][i] == (ga.doubles[1][ka[k]] if k in ka else -np.inf)
        assert g.longs[1][i] == (gb.longs[0][kb[k]] if k in kb else 0)


def test_filtered_aggregations_vs_numpy():
    """Independent restatement: decode the columns with numpy and compute the swim-lanes directly."""
    from pinot_b200 import datagen
    from pinot_b200.segment_writer import unpack_bits_be
    seg = datagen.make_segment_synth(3, 120_011, columns=["c1", "c2", "d1", "m0", "m1"])   # > METRIC_CARD rows: bit-packed (unsorted) metric columns

    def vals(name):
        c = seg.columns[name]
        return c.dictionary_values()[unpack_bits_be(c.forward_index, c.num_docs, c.bits_per_element)].astype(np.int64)

    c1, c2, d1, m0, m1 = (vals(n) for n in ("c1", "c2", "d1", "m0", "m1"))
    k1, k2 = int(np.median(c1)), int(np.percentile(c2, 30))
    q = parse_sql(f"SELECT d1, SUM(m0) FILTER(WHERE c2 < {k2}), COUNT(*) FILTER(WHERE c2 < {k2}), MAX(m1) FILTER(WHERE m0 > 900000), "
                  f"AVG(m1), MIN(m0) FILTER(WHERE c2 >= {k2} AND m1 < 1000) FROM t WHERE c1 > {k1} GROUP BY d1 LIMIT 1000")
    r = oracle.execute(seg, q)
    main = c1 > k1
    f1, f2, f3 = main & (c2 < k2), main & (m0 > 900000), main & (c2 >= k2) & (m1 < 1000)
    keys = {k[0]: i for i, k in enumerate(r.decoded_keys())}
    assert sorted(keys) == sorted(np.unique(d1[main]).tolist())
    for key, i in keys.items():
        g = d1 == key
        assert r.doubles[0][i] == float(m0[g & f1].sum()) and r.longs[1][i] == int((g & f1).sum())
        assert r.doubles[2][i] == (float(m1[g & f2].max()) if (g & f2).any() else -np.inf)
        assert r.doubles[3][i] == float(m1[g & main].sum()) and r.longs[3][i] == int((g & main).sum())
        assert r.doubles[4][i] == (float(m0[g & f3].min()) if (g & f3).any() else np.inf)
    # ExecutionStatistics lane by lane: three clause lanes + the non-filtered lane (FilteredGroupByOperator.java:146-149)
    assert r.stats["num_docs_scanned"] == int(f1.sum() + f2.sum() + f3.sum() + main.sum())
    assert r.stats["num_entries_scanned_post_filter"] == int(f1.sum() * 2 + f2.sum() * 2 + f3.sum() * 2 + main.sum() * 2)


# InterSegmentGroupBySingleValueQueriesTest.java:61-288: 4 identical segments merged by GroupByCombineOperator; the ORDER BY /
# LIMIT of those queries is broker-side and out of scope, the (group -> value) literals are not
G11_SUM1 = {b"": 5935285005452.0, b"P": 88832999206836.0, b"gFuH": 63202785888.0, b"o": 18105331533948.0, b"t": 16331923219264.0}
G11_12_SUM1 = {(b"", b"HEuxNvH"): 3789390396216.0, (b"", b"KrNxpdycSiwoRohEiTIlLqDHnx"): 733802350944.0,
               (b"", b"MaztCmmxxgguBUxPti"): 1333941430664.0, (b"", b"dJWwFk"): 55470665124.0, (b"", b"oZgnrlDEtjjVpUoFLol"): 22680162504.0,
               (b"P", b"HEuxNvH"): 21998672845052.0, (b"P", b"KrNxpdycSiwoRohEiTIlLqDHnx"): 18069909216728.0,
               (b"P", b"MaztCmmxxgguBUxPti"): 27177029040008.0, (b"P", b"TTltMtFiRqUjvOG"): 4462670055540.0, (b"P", b"XcBNHe"): 120021767504.0,
               (b"P", b"dJWwFk"): 6224665921376.0, (b"P", b"fykKFqiw"): 1574451324140.0, (b"P", b"gFuH"): 860077643636.0,
               (b"P", b"oZgnrlDEtjjVpUoFLol"): 8345501392852.0, (b"gFuH", b"HEuxNvH"): 29872400856.0,
               (b"o", b"MaztCmmxxgguBUxPti"): 6905624581072.0, (b"o", b"HEuxNvH"): 5026384681784.0, (b"t", b"MaztCmmxxgguBUxPti"): 4492405624940.0,
               (b"t", b"HEuxNvH"): 4424489490364.0, (b"o", b"KrNxpdycSiwoRohEiTIlLqDHnx"): 4051812250524.0,
               (b"t", b"KrNxpdycSiwoRohEiTIlLqDHnx"): 3529048341192.0, (b"t", b"dJWwFk"): 1349058948804.0, (b"o", b"dJWwFk"): 1152689463360.0,
               (b"t", b"oZgnrlDEtjjVpUoFLol"): 1039101333316.0, (b"o", b"oZgnrlDEtjjVpUoFLol"): 699381633640.0,
               (b"t", b"TTltMtFiRqUjvOG"): 675238030848.0, (b"t", b"fykKFqiw"): 480973878052.0, (b"t", b"gFuH"): 330331507792.0,
               (b"o", b"TTltMtFiRqUjvOG"): 203835153352.0, (b"o", b"fykKFqiw"): 62975165296.0, (b"gFuH", b"MaztCmmxxgguBUxPti"): 29170832184.0,
               (b"t", b"XcBNHe"): 11276063956.0, (b"gFuH", b"KrNxpdycSiwoRohEiTIlLqDHnx"): 4159552848.0, (b"o", b"gFuH"): 2628604920.0}
G11_MIN6 = {b"": 296467636.0, b"P": 1689277.0, b"gFuH": 296467636.0, b"o": 296467636.0, b"t": 1980174.0}
G12_MIN6 = {b"XcBNHe": 329467557.0, b"fykKFqiw": 296467636.0, b"gFuH": 296467636.0, b"HEuxNvH": 6043515.0, b"MaztCmmxxgguBUxPti": 6043515.0,
            b"dJWwFk": 6043515.0, b"KrNxpdycSiwoRohEiTIlLqDHnx": 1980174.0, b"TTltMtFiRqUjvOG": 1980174.0, b"oZgnrlDEtjjVpUoFLol": 1689277.0}
G17_COUNT = {83386499: 2924, 217787432: 3892, 227908817: 6564, 402773817: 7304, 423049234: 6556, 561673250: 7420, 635942547: 3308,
             638936844: 3816, 939479517: 3116, 984091268: 3824, 1230252339: 5620, 1284373442: 7428, 1555255521: 2900, 1618904660: 2744,
             1670085862: 3388}
G11_AVG6 = {b"": 296467636.0, b"P": 909380310.3521485, b"gFuH": 296467636.0, b"o": 296467636.0, b"t": 526245333.3900426}
G12_DC11 = {b"HEuxNvH": 5, b"KrNxpdycSiwoRohEiTIlLqDHnx": 5, b"MaztCmmxxgguBUxPti": 5, b"TTltMtFiRqUjvOG": 3, b"XcBNHe": 2, b"dJWwFk": 4,
            b"fykKFqiw": 3, b"gFuH": 3, b"oZgnrlDEtjjVpUoFLol": 4}


def test_inter_segment_group_by():
    t, _ = _inter("SELECT column11, SUM(column1), MIN(column6) FROM testTable GROUP BY column11")
    assert {k[0]: v[0] for k, v in t.items()} == G11_SUM1 and {k[0]: v[1] for k, v in t.items()} == G11_MIN6
    t, _ = _inter("SELECT column11, column12, SUM(column1) FROM testTable GROUP BY column11, column12")
    got = {k: v[0] for k, v in t.items()}
    assert all(got[k] == v for k, v in G11_12_SUM1.items())
    assert sorted(got.values(), reverse=True)[:len(G11_12_SUM1)] == sorted(G11_12_SUM1.values(), reverse=True)     # the top of ORDER BY SUM DESC LIMIT 50
    t, _ = _inter("SELECT column12, MIN(column6) FROM testTable GROUP BY column12")
    assert {k[0]: v[0] for k, v in t.items()} == G12_MIN6
    t, _ = _inter("SELECT column17, COUNT(*) FROM testTable GROUP BY column17")
    got = {k[0]: v[0] for k, v in t.items()}
    assert [(k, got[k]) for k in sorted(got)[:15]] == sorted(G17_COUNT.items())
    t, _ = _inter("SELECT column11, AVG(column6) FROM testTable GROUP BY column11")
    assert {k[0]: v[0][0] / v[0][1] for k, v in t.items()} == G11_AVG6
    t, _ = _inter("SELECT column12, DISTINCTCOUNT(column11) FROM testTable GROUP BY column12")
    assert {k[0]: len(v[0]) for k, v in t.items()} == G12_DC11
