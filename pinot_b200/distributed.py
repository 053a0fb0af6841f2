"""Multi-GPU bootstrap (one process per GPU).  Everything on the data path lives in libpinot_b200.so: segments shard
across ranks with no data-path collective, and the one exchange step -- the merge of the per-rank group tables, the
device-side equivalent of GroupByCombineOperator's IndexedTable merge (CTR/operator/combine/GroupByCombineOperator.java:
132-147) -- is done by pb_query_execute(PB_Q_ALL_RANKS) itself over NCCL.  What is left for the host language is the
rendezvous a JVM would do over its own control plane:

  1. init_comm: rank 0 makes the NCCL unique id (pb_comm_unique_id), everybody receives it and calls pb_comm_init;
  2. agree_global_dictionaries: every rank exports the sorted union of its segments' dictionaries for each group-by /
     DISTINCTCOUNT column, the unions are exchanged, and the union over all ranks is installed, so the dense tables (and
     distinct bitsets) of all ranks line up slot for slot.

The exchange itself is pluggable: `TorchExchange` (torch.distributed, any backend -- bench.py under torchrun) or
`FileExchange` (a shared directory; no torch involved -- tests, and the shape of what a JVM would do).
"""
from __future__ import annotations

import os
import pickle
import time
from typing import List, Sequence

import numpy as np

from . import native


class TorchExchange:
    def __init__(self, dist):
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def all_gather(self, obj) -> list:
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def barrier(self):
        self.dist.barrier()


class FileExchange:
    """all-gather of small python objects through files in a directory every rank can see"""

    def __init__(self, directory: str, rank: int, world: int, timeout_s: float = 120.0):
        self.dir, self.rank, self.world, self.timeout = directory, rank, world, timeout_s
        self.round = 0
        os.makedirs(directory, exist_ok=True)

    def all_gather(self, obj) -> list:
        r = self.round
        self.round += 1
        tmp = os.path.join(self.dir, f".x{r}_{self.rank}.tmp")
        with open(tmp, "wb") as f:
            pickle.dump(obj, f)
        os.replace(tmp, os.path.join(self.dir, f"x{r}_{self.rank}.pkl"))
        out, deadline = [], time.time() + self.timeout
        for k in range(self.world):
            path = os.path.join(self.dir, f"x{r}_{k}.pkl")
            while not os.path.exists(path):
                if time.time() > deadline:
                    raise TimeoutError(f"rank {self.rank}: rank {k} never wrote round {r}")
                time.sleep(0.002)
            with open(path, "rb") as f:
                out.append(pickle.load(f))
        return out

    def barrier(self):
        self.all_gather(None)


def init_comm(exchange) -> None:
    """ncclCommInitRank inside the library (collective)."""
    uid = native.comm_unique_id() if exchange.rank == 0 else None
    uid = exchange.all_gather(uid)[0]
    native.comm_init(exchange.world, exchange.rank, uid)


def merge_sorted_dictionaries(blocks: Sequence[np.ndarray], stored_type: int) -> np.ndarray:
    """blocks: [n_i, entry_bytes] uint8 arrays of native-endian entries -> sorted unique union (same layout)."""
    eb = max(b.shape[1] for b in blocks)
    if stored_type == 4:          # STRING: fixed-width zero-padded entries, ordered like String.compareTo (UTF-16 code units)
        from .segment_writer import java_string_key
        keys = set()
        for b in blocks:
            for r in np.ascontiguousarray(b, dtype=np.uint8):
                keys.add(bytes(r).ljust(eb, b"\0"))
        uniq = sorted(keys, key=lambda e: java_string_key(e.rstrip(b"\0")))
        return np.frombuffer(b"".join(uniq), dtype=np.uint8).reshape(len(uniq), eb).copy()
    allv = np.concatenate([np.ascontiguousarray(b, dtype=np.uint8) for b in blocks], axis=0)
    dt = {0: np.int32, 1: np.int64, 2: np.float32, 3: np.float64}[stored_type]
    vals = np.unique(allv.reshape(-1).view(dt))
    return vals.astype(dt).view(np.uint8).reshape(-1, eb).copy()


def agree_global_dictionaries(group: native.SegmentGroup, columns: Sequence[str], stored_types: Sequence[int], exchange) -> None:
    """exchange: a TorchExchange / FileExchange, or a torch.distributed module (wrapped)."""
    if not hasattr(exchange, "all_gather"):
        exchange = TorchExchange(exchange)
    for col, ty in zip(columns, stored_types):
        mine = group.export_dictionary(col)
        group.set_global_dictionary(col, merge_sorted_dictionaries(exchange.all_gather(mine), ty))


def dictionary_columns(q, segment) -> List[str]:
    """the columns whose global dictionaries the ranks must agree on for `q`: dictionary-encoded group-by columns and
    DISTINCTCOUNT inputs"""
    from .query import AggOp
    cols = [c for c in q.group_by if segment.columns[c].has_dictionary]
    for a in q.aggregations:
        if a.op == AggOp.DISTINCTCOUNT and a.column and segment.columns[a.column].has_dictionary and a.column not in cols:
            cols.append(a.column)
    return cols


def shard_segments(n_segments: int, rank: int, world: int) -> List[int]:
    """static round-robin of a table's segments over the ranks"""
    return list(range(rank, n_segments, world))
