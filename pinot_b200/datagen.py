"""Synthetic table generator (SURVEY.md §8d / BASELINE.md §3): 20 columns per segment, seed 42.

    c0..c7   INT dims, dictionary, unsorted; cardinalities 100000, 1000, 10000, 64, 16, 8, 4, 2   (c1, c3 inverted)
    d0..d4   INT dims, dictionary; cardinalities 8, 16, 32, 4, 4                                   (d0 inverted)
    s0       STRING dim, dictionary, cardinality 10000, 8 chars                                     (inverted)
    t0       INT, sorted (docId-monotone), cardinality rows/1000
    m0..m2   INT metrics, dictionary, cardinality 100000, values in [0, 10^6)  (sums < 2^53: bit-exact checks)
    x0, x1   DOUBLE metrics, raw PASS_THROUGH, uniform [0,1)
    k0       LONG, raw PASS_THROUGH, ~10M distinct values (config 5 key)

Every segment draws its own dictionaries (numpy Generator(PCG64(seed + segment_index))), so dictIds are
segment-local exactly as in a real table.  Emits Pinot-layout buffers through segment_writer.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .segment_writer import ColumnIndex, DataType, Segment, build_column, build_dict_column, make_segment

DIM_CARDS = {"c0": 100_000, "c1": 1_000, "c2": 10_000, "c3": 64, "c4": 16, "c5": 8, "c6": 4, "c7": 2,
             "d0": 8, "d1": 16, "d2": 32, "d3": 4, "d4": 4}
METRIC_CARD = 100_000
INVERTED = {"c1", "c3", "d0", "s0"}
ALL_COLUMNS = ["c0", "c1", "c2", "c3", "c4", "c5", "c6", "c7", "d0", "d1", "d2", "d3", "d4", "s0", "t0",
               "m0", "m1", "m2", "x0", "x1", "k0"]


def _ids(rng, card: int, n: int) -> np.ndarray:
    ids = rng.integers(0, card, size=n, dtype=np.uint32)
    k = min(card, n)
    # make sure every dictId occurs (Pinot dictionaries only hold values that are present)
    pos = rng.choice(n, size=k, replace=False) if n > k else np.arange(k)
    ids[pos] = np.arange(k, dtype=np.uint32)
    return ids


def _int_dictionary(rng, card: int, value_range: int) -> np.ndarray:
    card = min(card, value_range)
    return np.sort(rng.choice(value_range, size=card, replace=False)).astype(np.int32)


def make_column(name: str, n: int, rng, inverted: Optional[bool] = None, dict_rng=None) -> ColumnIndex:
    """dict_rng draws the dictionary of a dimension column (shared by all segments of a table unless the caller
    passes the per-segment stream, see make_segment_synth(vary_dim_dictionaries=True))."""
    inv = (name in INVERTED) if inverted is None else inverted
    if dict_rng is None:
        dict_rng = rng
    if name in DIM_CARDS:
        card = min(DIM_CARDS[name], n)
        dvals = _int_dictionary(dict_rng, card, max(card * 10, 1000))
        return build_dict_column(name, DataType.INT, dvals, _ids(rng, card, n), inverted=inv)
    if name in ("m0", "m1", "m2"):
        card = min(METRIC_CARD, n)
        dvals = _int_dictionary(rng, card, 1_000_000)
        return build_dict_column(name, DataType.INT, dvals, _ids(rng, card, n), inverted=False)
    if name == "s0":
        card = min(10_000, n)
        raw = dict_rng.choice(26 ** 4, size=card, replace=False)
        letters = np.array(list(b"abcdefghijklmnopqrstuvwxyz"), dtype=np.uint8)
        strs = []
        for v in np.sort(raw):
            a, b, c, d = (v // 17576) % 26, (v // 676) % 26, (v // 26) % 26, v % 26
            strs.append(bytes([letters[a], letters[b], letters[c], letters[d]]) + b"_key")
        dvals = np.array(sorted(strs), dtype=object)
        return build_dict_column(name, DataType.STRING, dvals, _ids(rng, card, n), inverted=inv)
    if name == "t0":
        card = max(1, n // 1000)
        ids = np.minimum(np.arange(n, dtype=np.int64) // 1000, card - 1).astype(np.uint32)
        dvals = (20_000 + np.arange(card)).astype(np.int32)
        return build_dict_column(name, DataType.INT, dvals, ids, inverted=False)
    if name in ("x0", "x1"):
        return build_column(name, DataType.DOUBLE, rng.random(n), dictionary=False)
    if name == "k0":
        vals = rng.integers(0, 10_000_000, size=n, dtype=np.int64) * 1_000_003 + 7
        return build_column(name, DataType.LONG, vals, dictionary=False)
    raise KeyError(name)


def make_segment_synth(index: int, num_docs: int, columns: Optional[Sequence[str]] = None, seed: int = 42,
                       name_prefix: str = "synth", vary_dim_dictionaries: bool = False) -> Segment:
    """Dimension dictionaries (c*, d*, s0) are table-wide, as dimension values are in a real table; metric
    dictionaries are per segment.  vary_dim_dictionaries=True draws the dimension dictionaries per segment too
    (exercises the local->global dictId remap of the combined mode)."""
    cols = list(columns) if columns is not None else ALL_COLUMNS
    out = []
    for cname in cols:
        ci = ALL_COLUMNS.index(cname)
        # one independent stream per (segment, column) so a column looks the same whichever subset is built
        rng = np.random.Generator(np.random.PCG64([seed + index, ci]))
        dict_rng = None if vary_dim_dictionaries else np.random.Generator(np.random.PCG64([seed, 1000 + ci]))
        out.append(make_column(cname, num_docs, rng, dict_rng=dict_rng))
    return make_segment(f"{name_prefix}_{index}", out)


def make_table(num_segments: int, docs_per_segment: int, columns: Optional[Sequence[str]] = None, seed: int = 42,
               first_index: int = 0, vary_dim_dictionaries: bool = False) -> List[Segment]:
    return [make_segment_synth(first_index + i, docs_per_segment, columns, seed, vary_dim_dictionaries=vary_dim_dictionaries)
            for i in range(num_segments)]


def _make_one(args):
    index, docs, columns, seed, vary = args
    return make_segment_synth(index, docs, columns, seed, vary_dim_dictionaries=vary)


def make_table_parallel(num_segments: int, docs_per_segment: int, columns: Optional[Sequence[str]] = None, seed: int = 42,
                        first_index: int = 0, vary_dim_dictionaries: bool = False, workers: int = 8,
                        indices: Optional[Sequence[int]] = None) -> List[Segment]:
    """make_table with one worker process per segment (spawned: safe after CUDA initialisation).  Every segment is drawn
    from its own (seed, index, column) streams, so the result is identical to the sequential generator's."""
    idx = list(indices) if indices is not None else [first_index + i for i in range(num_segments)]
    if workers <= 1 or len(idx) <= 1:
        return [_make_one((i, docs_per_segment, columns, seed, vary_dim_dictionaries)) for i in idx]
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    with ProcessPoolExecutor(max_workers=min(workers, len(idx)), mp_context=mp.get_context("spawn")) as ex:
        return list(ex.map(_make_one, [(i, docs_per_segment, columns, seed, vary_dim_dictionaries) for i in idx]))


# ---- BASELINE.json configs as SQL (literals are chosen per table, see helpers) ----

def config1_sql(seg: Segment) -> str:
    """SELECT SUM(c0) WHERE c1 > k, k = dictionary value at 50 %."""
    d = seg.columns["c1"].dictionary_values()
    return f"SELECT SUM(c0) FROM t WHERE c1 > {int(d[len(d) // 2 - 1])}"


def config2_sql(segs: Sequence[Segment], in_values: int = 16) -> str:
    """WHERE c1 IN (n values) AND c2 < k(50 %) GROUP BY d0,d1,d2 -> SUM(m0), COUNT(*), MIN(m1), MAX(m2);
    inverted index on c1 disabled via skipIndexes so both predicates scan."""
    d1 = segs[0].columns["c1"].dictionary_values()
    step = max(1, len(d1) // in_values)
    vals = [int(v) for v in d1[::step][:in_values]]
    d2 = segs[0].columns["c2"].dictionary_values()
    k = int(d2[len(d2) // 2])
    return ("SET skipIndexes = 'c1=inverted'; SELECT d0, d1, d2, SUM(m0), COUNT(*), MIN(m1), MAX(m2) FROM t "
            f"WHERE c1 IN ({', '.join(map(str, vals))}) AND c2 < {k} GROUP BY d0, d1, d2 LIMIT 100000")


CONFIG2_COLUMNS = ["c1", "c2", "d0", "d1", "d2", "m0", "m1", "m2"]
CONFIG2_BITS_PER_ROW = 10 + 14 + 3 + 4 + 5 + 17 + 17 + 17      # = 87
