"""Generate tests/golden/test_data_sv.npz from the reference's own test fixture.

Run ONCE in the build container (it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Source: /root/reference/pinot-core/src/test/resources/data/test_data-sv.avro (30 000 rows, 18 columns,
null codec).  The 11 columns kept are the ones BaseSingleValueQueriesTest.java:51-85 builds its segment
from.  The golden aggregates checked against this data are literal values quoted from
InnerSegmentAggregationSingleValueQueriesTest.java / InterSegmentAggregationSingleValueQueriesTest.java
(see tests/test_oracle_golden.py for file:line of each).
"""
import json
import os
import sys

import numpy as np

SRC = "/root/reference/pinot-core/src/test/resources/data/test_data-sv.avro"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_data_sv.npz")
KEEP = ["column1", "column3", "column5", "column6", "column7", "column9", "column11", "column12",
        "column17", "column18", "daysSinceEpoch"]


def read_avro(path):
    b = open(path, "rb").read()
    assert b[:4] == b"Obj\x01"
    pos = 4

    def rl():
        nonlocal pos
        r = 0
        s = 0
        while True:
            c = b[pos]
            pos += 1
            r |= (c & 0x7F) << s
            s += 7
            if not c & 0x80:
                break
        return (r >> 1) ^ -(r & 1)

    meta = {}
    while True:
        n = rl()
        if n == 0:
            break
        if n < 0:
            n = -n
            rl()
        for _ in range(n):
            kl = rl()
            k = b[pos:pos + kl]
            pos += kl
            vl = rl()
            meta[k] = b[pos:pos + vl]
            pos += vl
    assert meta.get(b"avro.codec", b"null") == b"null"
    schema = json.loads(meta[b"avro.schema"])
    fields = [(f["name"], [t for t in f["type"] if t != "null"][0]) for f in schema["fields"]]
    sync = b[pos:pos + 16]
    pos += 16
    cols = {name: [] for name, _ in fields}
    while pos < len(b):
        count = rl()
        rl()  # block byte size
        for _ in range(count):
            for name, typ in fields:
                branch = rl()          # union index: 0 = null, 1 = value
                if branch == 0:
                    cols[name].append(None)
                elif typ == "int":
                    cols[name].append(rl())
                elif typ == "string":
                    ln = rl()
                    cols[name].append(b[pos:pos + ln])
                    pos += ln
                else:
                    raise ValueError(typ)
        assert b[pos:pos + 16] == sync
        pos += 16
    return fields, cols


def main():
    fields, cols = read_avro(SRC)
    types = dict(fields)
    out = {}
    for name in KEEP:
        v = cols[name]
        assert all(x is not None for x in v), name
        if types[name] == "int":
            out[name] = np.asarray(v, dtype=np.int32)
        else:
            out[name] = np.asarray(v, dtype="S")
    n = len(out["column1"])
    assert n == 30000
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", n, "rows")


if __name__ == "__main__":
    sys.exit(main())
