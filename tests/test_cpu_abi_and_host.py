"""CPU-only checks: the C-ABI library loads and exports every symbol the headers declare; the host planning layer
(FilterPlanNode / PredicateEvaluator / FilterOperatorUtils restatement) lowers filters as the reference does.  No
compute call is made here (there is no GPU)."""
import os
import re

import numpy as np
import pytest

from pinot_b200 import datagen, native
from pinot_b200.query import parse_sql
from pinot_b200.segment_writer import DataType, build_column, make_segment
from tests.fixtures import FILTER, sv_segment

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pbh?_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = native.lib()
    names = _declared("pinot_b200.h") + _declared("pinot_b200_host.h")
    assert len(names) > 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"symbols declared in include/*.h but not exported: {missing}"


def test_errors_are_reported_not_fatal():
    lib = native.lib()
    assert lib.pb_result_num_tables(None) == 0
    assert lib.pb_segment_release(None) == 0
    assert native.PinotB200Error(-2, "x").code == -2


@pytest.fixture(scope="module")
def sv_group():
    seg = sv_segment()
    st = native.StagedSegment(seg)       # registers buffers only: no device needed for planning
    g = native.SegmentGroup([st])
    yield seg, g
    g.release()


def test_filter_plan_matches_reference_operator_selection(sv_group):
    """BaseSingleValueQueriesTest FILTER: column5='gFuH' matches all (cardinality 1) and is dropped; daysSinceEpoch is
    sorted -> sorted index; column11 NOT IN uses the inverted index; column1/column3/column6 ranges scan; AND children
    ordered sorted < OR < scans (FilterOperatorUtils.java:205-252)."""
    seg, g = sv_group
    q = parse_sql("SELECT COUNT(*) FROM testTable" + FILTER)
    plan = native.explain_filter(g, q).splitlines()
    assert plan[0] == "FILTER_AND"
    kinds = [ln.strip().split("(")[0] for ln in plan[1:]]
    assert kinds == ["FILTER_SORTED_INDEX", "FILTER_OR", "FILTER_FULL_SCAN", "FILTER_INVERTED_INDEX", "FILTER_FULL_SCAN", "FILTER_FULL_SCAN"]
    assert "column6 RANGE" in plan[3] and "column11 NOT_IN" in plan[4]
    assert "column1 RANGE" in plan[5] and "column3 RANGE" in plan[6]
    assert native.is_eligible(g, q)


def test_predicate_lowering_edge_cases(sv_group):
    seg, g = sv_group
    # value absent from the dictionary: EQ -> empty, NOT_EQ -> match all
    assert native.explain_filter(g, parse_sql("SELECT COUNT(*) FROM t WHERE column1 = 1")).strip() == "FILTER_EMPTY"
    assert native.explain_filter(g, parse_sql("SELECT COUNT(*) FROM t WHERE column1 <> 1")).strip() == "FILTER_MATCH_ENTIRE_SEGMENT"
    # range covering the whole dictionary -> match all; AND with an empty child -> empty; OR with match-all -> match all
    assert native.explain_filter(g, parse_sql("SELECT COUNT(*) FROM t WHERE column1 > 0")).strip() == "FILTER_MATCH_ENTIRE_SEGMENT"
    assert native.explain_filter(g, parse_sql("SELECT COUNT(*) FROM t WHERE column1 = 1 AND column3 > 5")).strip() == "FILTER_EMPTY"
    assert native.explain_filter(g, parse_sql("SELECT COUNT(*) FROM t WHERE column1 > 0 OR column3 = 5")).strip() == "FILTER_MATCH_ENTIRE_SEGMENT"
    # NOT of an inverted-index leaf keeps the index; skipIndexes forces a scan
    assert "FILTER_INVERTED_INDEX" in native.explain_filter(g, parse_sql("SELECT COUNT(*) FROM t WHERE NOT column7 = 296467636"))
    assert "FILTER_FULL_SCAN" in native.explain_filter(g, parse_sql("SET skipIndexes='column7=inverted'; SELECT COUNT(*) FROM t WHERE column7 = 296467636"))
    # dictId range of a sorted-dictionary RANGE equals numpy's searchsorted on the dictionary
    d = seg.columns["column3"].dictionary_values()
    txt = native.explain_filter(g, parse_sql("SELECT COUNT(*) FROM t WHERE column3 BETWEEN 20000000 AND 1000000000"))
    lo, hi = int(np.searchsorted(d, 20000000, "left")), int(np.searchsorted(d, 1000000000, "right"))
    assert f"dictIds[{lo},{hi})" in txt


def test_filtered_aggregation_clauses_are_planned_on_their_own(sv_group):
    """AggregationFunctionUtils.buildFilteredAggregationInfos (:312-400): every FILTER clause gets a FilterPlanNode of its own;
    IS [NOT] NULL without a null-value vector is EmptyFilterOperator / MatchAllFilterOperator (FilterPlanNode.java:294-307)."""
    seg, g = sv_group
    q = parse_sql("SELECT SUM(column6) FILTER(WHERE column6 > 5 OR column6 < 15), COUNT(*) FILTER(WHERE column1 IS NOT NULL), "
                  "MAX(column3) FILTER(WHERE column3 IS NULL), SUM(column3), AVG(column7) FILTER(WHERE column7 IN (296467636, 1111197135) AND column3 > 100000000), "
                  "COUNT(*) FILTER(WHERE column1 IS NOT NULL) FROM testTable WHERE daysSinceEpoch > 126164076")
    filters, index = q.agg_filters()
    assert index == [0, 1, 2, -1, 3, 1] and len(filters) == 4          # equal clauses share a swim-lane
    assert native.is_eligible(g, q)
    assert native.explain_agg_filter(g, q, 0).strip() == "FILTER_MATCH_ENTIRE_SEGMENT"
    assert native.explain_agg_filter(g, q, 1).strip() == "FILTER_MATCH_ENTIRE_SEGMENT"
    assert native.explain_agg_filter(g, q, 2).strip() == "FILTER_EMPTY"
    txt = native.explain_agg_filter(g, q, 3)
    assert txt.splitlines()[0] == "FILTER_AND" and "FILTER_INVERTED_INDEX(column7 IN n=2)" in txt and "FILTER_FULL_SCAN(column3 RANGE" in txt
    assert "FILTER_SORTED_INDEX" in native.explain_filter(g, q)        # the main filter is unaffected
    assert not native.is_eligible(g, parse_sql("SELECT COUNT(*) FILTER(WHERE nosuch > 3) FROM testTable"))


def test_ineligible_queries_decline(sv_group):
    seg, g = sv_group
    assert not native.is_eligible(g, parse_sql("SELECT SUM(column11) FROM t"))          # numeric aggregation on STRING
    assert not native.is_eligible(g, parse_sql("SELECT COUNT(*) FROM t WHERE nosuch = 3"))


def test_global_dictionary_remaps_on_host():
    segs = [datagen.make_segment_synth(i, 5000, columns=["d1", "m0"], vary_dim_dictionaries=True) for i in range(3)]
    staged = [native.StagedSegment(s) for s in segs]
    g = native.SegmentGroup(staged)
    union = g.export_dictionary("d1").view(np.int32).reshape(-1)
    exp = np.unique(np.concatenate([s.columns["d1"].dictionary_values() for s in segs]))
    assert (union == exp).all()
    for i, s in enumerate(segs):
        rm = g.remap("d1", i)
        assert (union[rm] == s.columns["d1"].dictionary_values()).all()
    g.release()


def test_stage_and_plan_from_an_mmapped_v3_directory(tmp_path):
    """pb_segment_stage + host planning over index buffers that are views of one mmap'd columns.psf (arbitrary byte offsets,
    read-only pages) — what SegmentDirectory.Reader.getIndexFor hands out in a server.  No device work happens here:
    columns are copied to HBM on first use by a query."""
    from pinot_b200.segment_writer import load_v3, write_v3
    seg = datagen.make_segment_synth(4, 30_011, columns=["c1", "c3", "d0", "s0", "t0", "m0", "x0", "k0"])
    back = load_v3(write_v3(seg, str(tmp_path)))
    g_mem = native.SegmentGroup([native.StagedSegment(seg)])
    g_map = native.SegmentGroup([native.StagedSegment(back)])
    d1, d3 = seg.columns["c1"].dictionary_values(), seg.columns["c3"].dictionary_values()
    for sql in (f"SELECT s0, COUNT(*), SUM(m0) FROM t WHERE c3 IN ({int(d3[2])}, {int(d3[9])}) AND c1 BETWEEN {int(d1[40])} AND {int(d1[700])} GROUP BY s0",
                "SELECT COUNT(*) FROM t WHERE t0 BETWEEN 20003 AND 20011 OR x0 < 0.25 OR k0 IN (7, 1000010)",
                f"SELECT MAX(k0) FILTER(WHERE s0 = 'aaaa_key' OR c1 > {int(d1[990])}) FROM t WHERE NOT c3 = {int(d3[1])}"):
        q = parse_sql(sql)
        assert native.is_eligible(g_map, q)
        assert native.explain_filter(g_map, q) == native.explain_filter(g_mem, q)
    q = parse_sql(f"SELECT MAX(k0) FILTER(WHERE s0 = 'aaaa_key' OR c1 > {int(d1[990])}) FROM t")
    assert native.explain_agg_filter(g_map, q, 0) == native.explain_agg_filter(g_mem, q, 0)
    assert np.array_equal(np.asarray(g_map.export_dictionary("s0")), np.asarray(g_mem.export_dictionary("s0")))
    g_mem.release()
    g_map.release()


def test_null_handling_is_lowered_to_trues_programs():
    """enableNullHandling: the host layer hands the device the TRUES of the three-valued operator tree (BaseFilterOperator.
    getTrues / getNulls / getFalses and the And / Or / Not / BaseColumnFilterOperator overrides) as an ordinary program, and
    the clause "<column> IS NOT NULL" for every aggregation over a nullable column."""
    from pinot_b200.segment_writer import with_nulls
    rng = np.random.default_rng(1)
    n = 200
    a = rng.integers(0, 5, n).astype(np.int32); an = rng.random(n) < 0.3
    seg = make_segment("nh", [with_nulls(build_column("a", DataType.INT, a), an), build_column("b", DataType.INT, rng.integers(0, 9, n).astype(np.int32))])
    g = native.SegmentGroup([native.StagedSegment(seg)])
    nh = "SET enableNullHandling=true; "
    low = lambda sql, clause=-1: native.dump_lowered(g, parse_sql(sql), clause)
    # column leaf: matches AND NOT nulls (BaseColumnFilterOperator.java:46-54)
    assert low(nh + "SELECT COUNT(*) FROM t WHERE a > 2") == ["BITMAP col=a excl=1 null_value_vector", "SCAN_DICT_RANGE col=a lo=3 hi=5", "AND n=2"]
    # NOT leaf: NOT (trues OR nulls) (BaseFilterOperator.java:97-113)
    assert low(nh + "SELECT COUNT(*) FROM t WHERE NOT (a > 2)") == ["BITMAP col=a excl=1 null_value_vector", "SCAN_DICT_RANGE col=a lo=3 hi=5", "AND n=2",
                                                                  "BITMAP col=a excl=0 null_value_vector", "OR n=2", "NOT"]
    # a column without a null-value vector plans as before; so does everything without the option
    assert low(nh + "SELECT COUNT(*) FROM t WHERE NOT (b > 2)") == low("SELECT COUNT(*) FROM t WHERE NOT (b > 2)")
    assert low("SELECT COUNT(*) FROM t WHERE NOT (a > 2)") == ["SCAN_DICT_RANGE col=a lo=3 hi=5", "NOT"]
    # an always-true predicate on a nullable column is the flipped null bitmap (FilterOperatorUtils.java:78-88)
    assert low(nh + "SELECT COUNT(*) FROM t WHERE a >= 0") == ["BITMAP col=a excl=1 null_value_vector"]
    # nullable group-by keys keep the CPU plan
    assert not native.is_eligible(g, parse_sql(nh + "SELECT a, COUNT(*) FROM t GROUP BY a LIMIT 10"))
    assert native.is_eligible(g, parse_sql("SELECT a, COUNT(*) FROM t GROUP BY a LIMIT 10"))
    assert native.is_eligible(g, parse_sql(nh + "SELECT b, SUM(a) FROM t GROUP BY b LIMIT 10"))
    g.release()


def test_signed_zero_and_nan_in_raw_float_predicates():
    """Raw FLOAT / DOUBLE columns: EQ / NOT_EQ compare with == / != (0.0 equals -0.0; EqualsPredicateEvaluatorFactory.java:336-337),
    IN / NOT_IN ask a fastutil DoubleSet, i.e. compare Double.doubleToLongBits (-0.0 is not in {0.0}; InPredicateEvaluatorFactory.
    java:341-362).  The device compares bit patterns, so the host layer hands EQ both zeros; the oracle follows the same rules."""
    import struct
    from oracle import oracle
    x = np.array([0.0, -0.0, 1.5, -1.5, 0.0, -0.0, 2.0], dtype=np.float64)
    seg = make_segment("z", [build_column("x", DataType.DOUBLE, x, dictionary=False), build_column("f", DataType.FLOAT, x.astype(np.float32), dictionary=False)])
    g = native.SegmentGroup([native.StagedSegment(seg)])
    pos, neg = struct.unpack("<q", struct.pack("<d", 0.0))[0], struct.unpack("<q", struct.pack("<d", -0.0))[0]
    cnt = lambda where: int(oracle.execute(seg, parse_sql(f"SELECT COUNT(*) FROM t WHERE {where}")).longs[0][0])
    low = lambda where: native.dump_lowered(g, parse_sql(f"SELECT COUNT(*) FROM t WHERE {where}"))
    for col in ("x", "f"):
        assert low(f"{col} = 0.0") == [f"SCAN_RAW_SET col={col} excl=0 vals={pos},{neg}"] and cnt(f"{col} = 0.0") == 4
        assert low(f"{col} = -0.0") == [f"SCAN_RAW_SET col={col} excl=0 vals={neg},{pos}"] and cnt(f"{col} = -0.0") == 4
        assert low(f"{col} <> 0.0") == [f"SCAN_RAW_SET col={col} excl=1 vals={pos},{neg}"] and cnt(f"{col} <> 0.0") == 3
        assert low(f"{col} IN (0.0, 1.5)")[0].startswith(f"SCAN_RAW_SET col={col} excl=0 vals={pos},") and cnt(f"{col} IN (0.0, 1.5)") == 3
        assert cnt(f"{col} NOT IN (0.0)") == 5 and cnt(f"{col} IN (-0.0)") == 2
    g.release()


def test_strings_compare_like_java_strings():
    """A STRING dictionary is sorted by String.compareTo -- UTF-16 code units -- and searched with ValueReaderComparisons.
    compareUtf8Bytes (SEGL/io/util/ValueReaderComparisons.java:68-139).  That differs from the byte order of the UTF-8
    encodings exactly between supplementary characters (surrogate pairs) and BMP characters from U+E000 up: a dictionary that
    holds both an emoji and a full-width letter.  Host layer, oracle, the writer and the global-dictionary merge of the
    engine must all use the Java order, or values are not found."""
    from oracle import oracle
    from pinot_b200.distributed import merge_sorted_dictionaries
    from pinot_b200.segment_writer import java_string_key
    words = ["apple", "zebra", "\U0001F600", "\U0001F600x", "Ａ", "Ａb", "", "中", "\U00020000", "z\U0001F600", "zＡ", ""]
    java_order = sorted(words, key=lambda w: w.encode("utf-16-be", "surrogatepass"))
    assert java_order != sorted(words, key=lambda w: w.encode("utf-8"))              # the two orders really differ here
    rng = np.random.default_rng(3)
    def seg(name, subset, n):
        vals = [subset[i] for i in rng.integers(0, len(subset), n)]
        c = build_column("s", DataType.STRING, vals)
        assert [v.decode() for v in c.dictionary_values()] == sorted(set(vals), key=lambda w: w.encode("utf-16-be", "surrogatepass"))
        return make_segment(name, [c, build_column("d", DataType.INT, rng.integers(0, 3, n).astype(np.int32))]), np.array(vals, dtype=object)
    s0, v0 = seg("u0", words, 3000)
    s1, v1 = seg("u1", words[2:9], 2000)
    g = native.SegmentGroup([native.StagedSegment(s0), native.StagedSegment(s1)])
    key = lambda w: w.encode("utf-16-be", "surrogatepass")
    cases = {f"s = '{w}'": (lambda v, w=w: v == w) for w in words if w}
    cases.update({f"s > '{w}'": (lambda v, w=w: np.array([key(x) > key(w) for x in v])) for w in ("Ａ", "\U0001F600", "zebra", "中")})
    cases.update({f"s <= '{w}'": (lambda v, w=w: np.array([key(x) <= key(w) for x in v])) for w in ("\U0001F600x", "")})
    cases["s IN ('\U0001F600', 'Ａ', 'nope')"] = lambda v: np.isin(v, ["\U0001F600", "Ａ"])
    cases["s BETWEEN '\U0001F600' AND 'Ａ'"] = lambda v: np.array([key("\U0001F600") <= key(x) <= key("Ａ") for x in v])
    from tests.test_cpu_lowering_fuzz import evaluate_lowered
    for si, (sg, v) in enumerate(((s0, v0), (s1, v1))):
        for where, f in cases.items():
            q = parse_sql("SELECT COUNT(*) FROM t WHERE " + where)
            exp = np.nonzero(np.asarray(f(v), dtype=bool))[0].tolist()
            assert oracle.filter_doc_ids(sg, q)[0].tolist() == exp, (si, where)
            assert np.nonzero(evaluate_lowered(sg, native.dump_lowered(g, q, -1, si)))[0].tolist() == exp, (si, where)
    # the engine's union of the two dictionaries and its local -> global remaps
    union = g.export_dictionary("s")
    got = [bytes(r).rstrip(b"\0").decode() for r in union]
    assert got == java_order
    for si, sg in enumerate((s0, s1)):
        local = [x.decode() for x in sg.columns["s"].dictionary_values()]
        assert [got[j] for j in g.remap("s", si)] == local
    merged = merge_sorted_dictionaries([union, union[2:5]], 4)
    assert [bytes(r).rstrip(b"\0").decode() for r in merged] == java_order
    g.set_global_dictionary("s", merged)
    g.release()
