"""BASELINE-size checks (-m gpu): a full 12.5 M-row segment of configs[1] against the oracle AND an independent numpy
group-by over the unpacked columns; plus size-independent properties (partition additivity, count conservation)."""
import numpy as np
import pytest

from oracle import oracle
from pinot_b200 import datagen, native
from pinot_b200.query import parse_sql
from pinot_b200.segment_writer import unpack_bits_be
from tests.parity import assert_rows_equal, combined_rows, oracle_rows

pytestmark = pytest.mark.gpu

N = 12_500_000


@pytest.fixture(scope="module")
def big():
    native.init()
    segs = [datagen.make_segment_synth(i, N, columns=datagen.CONFIG2_COLUMNS) for i in range(2)]
    staged = [native.StagedSegment(s) for s in segs]
    g = native.SegmentGroup(staged)
    yield segs, g
    g.release()


def _values(seg, name):
    c = seg.columns[name]
    ids = unpack_bits_be(c.forward_index, c.num_docs, c.bits_per_element)
    return c.dictionary_values()[ids]


@pytest.mark.parametrize("in_values", [16, 500])
def test_config2_full_segment_vs_oracle_and_numpy(big, in_values):
    segs, g = big
    sql = datagen.config2_sql(segs, in_values)
    q = parse_sql(sql)
    res = native.execute(g, q)
    for i, seg in enumerate(segs[:1]):
        o = oracle.execute(seg, q)
        assert_rows_equal(res.tables[i].rows(), oracle_rows(o), q, exact_float=True, what=f"segment {i}")
        assert res.tables[i].stats["num_docs_scanned"] == o.stats["num_docs_scanned"]
        # independent numpy restatement of the same query
        preds = q.filter_postfix()[1]
        c1, c2 = _values(seg, "c1"), _values(seg, "c2")
        m = np.isin(c1, np.array([int(v) for v in preds[0].values], dtype=c1.dtype)) & (c2 < int(preds[1].upper))
        d = [_values(seg, n)[m].astype(np.int64) for n in ("d0", "d1", "d2")]
        key = (d[0] * 100_000 + d[1]) * 100_000 + d[2]
        uk, inv = np.unique(key, return_inverse=True)
        cnt = np.bincount(inv)
        sm = np.bincount(inv, weights=_values(seg, "m0")[m].astype(np.float64))
        rows = res.tables[i].rows()
        assert len(rows) == len(uk) and int(m.sum()) == res.tables[i].stats["num_docs_scanned"]
        got = {(k[0] * 100_000 + k[1]) * 100_000 + k[2]: v for k, v in rows.items()}
        for k, c, s in zip(uk.tolist(), cnt.tolist(), sm.tolist()):
            assert got[k][1] == c and got[k][0] == s


def test_properties_at_full_size(big):
    segs, g = big
    q = parse_sql(datagen.config2_sql(segs, 500))
    per = native.execute(g, q)
    comb = native.execute(g, q, native.PB_Q_COMBINE)
    # count conservation: sum of group counts == docs that passed the filter
    for t in per.tables + comb.tables:
        assert int(np.sum(t.longs[1])) == t.stats["num_docs_scanned"]
    # partition additivity: merged table == merge of the per-segment tables (dimension dictionaries are table-wide)
    merged = {}
    for t in per.tables:
        for k, row in t.rows().items():
            cur = merged.get(k)
            merged[k] = row if cur is None else [cur[0] + row[0], cur[1] + row[1], min(cur[2], row[2]), max(cur[3], row[3])]
    assert_rows_equal(comb.tables[0].rows(), merged, q, exact_float=True, what="partition additivity")
    # idempotence: same call, same answer
    again = native.execute(g, q, native.PB_Q_COMBINE)
    assert again.tables[0].rows() == comb.tables[0].rows()
    # unfiltered COUNT(*) == numTotalDocs
    c = native.execute(g, parse_sql("SELECT COUNT(*) FROM t"), native.PB_Q_COMBINE)
    assert c.tables[0].rows()[()][0] == 2 * N
