"""Multi-GPU plumbing (one process per GPU, torch.distributed): segments shard across ranks with no data-path
collective; the only exchange step is the merge of the per-rank group tables — the device-side equivalent of
GroupByCombineOperator's IndexedTable merge (CTR/operator/combine/GroupByCombineOperator.java:132-147).

  1. agree_global_dictionaries: every rank exports the sorted union of its segments' dictionaries for each
     group-by column, all-gathers them, and installs the union over all ranks, so dense tables line up.
  2. all_reduce_tables: in-place NCCL all-reduce of the dense table arrays left on the device by
     PB_Q_COMBINE | PB_Q_DEFER_FINALIZE (row counts + sums: SUM; min/max: MIN — MAX tables hold complements).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from . import native
from .query import AggOp, QueryContext


def merge_sorted_dictionaries(blocks: Sequence[np.ndarray], stored_type: int) -> np.ndarray:
    """blocks: [n_i, entry_bytes] uint8 arrays of native-endian entries -> sorted unique union (same layout)."""
    eb = blocks[0].shape[1]
    allv = np.concatenate([np.ascontiguousarray(b, dtype=np.uint8) for b in blocks], axis=0)
    if stored_type == 4:          # STRING: fixed-width padded entries compare bytewise
        keys = [bytes(r) for r in allv]
        uniq = sorted(set(keys))
        return np.frombuffer(b"".join(uniq), dtype=np.uint8).reshape(len(uniq), eb).copy()
    dt = {0: np.int32, 1: np.int64, 2: np.float32, 3: np.float64}[stored_type]
    vals = np.unique(allv.reshape(-1).view(dt))
    return vals.astype(dt).view(np.uint8).reshape(-1, eb).copy()


def agree_global_dictionaries(group: native.SegmentGroup, columns: Sequence[str], stored_types: Sequence[int], dist) -> None:
    for col, ty in zip(columns, stored_types):
        mine = group.export_dictionary(col)
        gathered = [None] * dist.get_world_size()
        dist.all_gather_object(gathered, mine)
        group.set_global_dictionary(col, merge_sorted_dictionaries(gathered, ty))


class _DevBuf:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def all_reduce_tables(result: native.Result, q: QueryContext, dist, torch) -> None:
    """NCCL all-reduce of a deferred, combined, dense result's device tables, in place, enqueued on the result's own
    CUDA stream (so it runs after the kernels and before pb_result_finalize without any host synchronisation).
    Three small collectives: counters + row counts (int64 SUM), sums (float64 SUM), min/max (int64 MIN)."""
    ext = torch.cuda.ExternalStream(result.stream())
    with torch.cuda.stream(ext):
        for which, typestr, op in ((5, "<i8", dist.ReduceOp.SUM), (6, "<f8", dist.ReduceOp.SUM), (7, "<i8", dist.ReduceOp.MIN)):
            ptr, n = result.device_buffer(which)
            if n > 0 and ptr:
                dist.all_reduce(torch.as_tensor(_DevBuf(ptr, n, typestr), device="cuda"), op=op)


_gather_buffers = {}


def all_gather_merge_tables(result: native.Result, dist, torch) -> None:
    """One collective instead of three: all-gather the whole table block of every rank on the result's stream, then
    pb_result_merge_gathered reduces the copies on the device (u64 SUM | f64 SUM | bitset OR | i64 MIN per region).
    The receive buffer is allocated once per size and reused (no allocator traffic on the hot path)."""
    ext = torch.cuda.ExternalStream(result.stream())
    ptr, nbytes = result.device_buffer(8)
    words, world = nbytes // 8, dist.get_world_size()
    gathered = _gather_buffers.get((words, world))
    if gathered is None:
        gathered = torch.empty(world * words, dtype=torch.int64, device="cuda")
        _gather_buffers[(words, world)] = gathered
    with torch.cuda.stream(ext):
        local = torch.as_tensor(_DevBuf(ptr, words, "<i8"), device="cuda")
        dist.all_gather_into_tensor(gathered, local)
        result.merge_gathered(gathered.data_ptr(), world)
