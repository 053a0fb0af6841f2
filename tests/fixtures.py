"""Shared fixtures: the reference's single-value test segment, rebuilt from tests/golden/test_data_sv.npz
with the schema / index config of BaseSingleValueQueriesTest.java:51-107 (inverted index on column6, column7,
column11, column17, column18; column5 and daysSinceEpoch come out sorted)."""
import functools
import os

import numpy as np

from pinot_b200.segment_writer import DataType, build_column, make_segment

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "test_data_sv.npz")

SV_SCHEMA = [("column1", DataType.INT), ("column3", DataType.INT), ("column5", DataType.STRING),
             ("column6", DataType.INT), ("column7", DataType.INT), ("column9", DataType.INT),
             ("column11", DataType.STRING), ("column12", DataType.STRING), ("column17", DataType.INT),
             ("column18", DataType.INT), ("daysSinceEpoch", DataType.INT)]
SV_INVERTED = {"column6", "column7", "column11", "column17", "column18"}

# BaseSingleValueQueriesTest.java:101-106
FILTER = (" WHERE column1 > 100000000"
          " AND column3 BETWEEN 20000000 AND 1000000000"
          " AND column5 = 'gFuH'"
          " AND (column6 < 500000000 OR column11 NOT IN ('t', 'P'))"
          " AND daysSinceEpoch = 126164076")


@functools.lru_cache(maxsize=None)
def sv_rows():
    d = np.load(GOLDEN)
    return {k: d[k] for k in d.files}


@functools.lru_cache(maxsize=None)
def sv_segment(name="testTable_126164076_167572854"):
    rows = sv_rows()
    cols = []
    for cname, dt in SV_SCHEMA:
        vals = rows[cname]
        if dt == DataType.STRING:
            vals = [bytes(v) for v in vals]
        cols.append(build_column(cname, dt, vals, dictionary=True, inverted=cname in SV_INVERTED))
    return make_segment(name, cols)
