"""World-size-2 gloo test (CPU) of the N>1 path's host logic: ranks own different segments with different
dictionaries, agree on global dictionaries, remap their local tables into the global key space and all-reduce
dense arrays; the result must equal the oracle's cross-segment merge over all segments."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle
from pinot_b200 import datagen, native
from pinot_b200.distributed import agree_global_dictionaries
from pinot_b200.query import parse_sql

SQL = "SELECT d0, d1, SUM(m0), COUNT(*), MIN(m1), MAX(m2) FROM t WHERE c2 < 60000 GROUP BY d0, d1 LIMIT 100000"
COLS = ["c2", "d0", "d1", "m0", "m1", "m2"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q = parse_sql(SQL)
    segs = [datagen.make_segment_synth(rank * 2 + i, 20_000, columns=COLS, vary_dim_dictionaries=True) for i in range(2)]
    staged = [native.StagedSegment(s) for s in segs]
    group = native.SegmentGroup(staged)
    agree_global_dictionaries(group, q.group_by, [0, 0], dist)
    gd = [group.export_dictionary(c).view(np.int32).reshape(-1) for c in q.group_by]
    cards = [len(d) for d in gd]
    G = cards[0] * cards[1]
    cnt = np.zeros(G, dtype=np.int64)
    sm = np.zeros(G)
    mn = np.full(G, np.inf)
    mx = np.full(G, -np.inf)
    for si, s in enumerate(segs):      # per-segment operator results -> global key space (what the device table holds)
        r = oracle.execute(s, q)
        rm0, rm1 = group.remap("d0", si), group.remap("d1", si)
        slot = rm0[r.group_keys[:, 0]] + cards[0] * rm1[r.group_keys[:, 1]]
        np.add.at(sm, slot, r.doubles[0]); np.add.at(cnt, slot, r.longs[1])
        np.minimum.at(mn, slot, r.doubles[2]); np.maximum.at(mx, slot, r.doubles[3])
    tc, ts, tmn, tmx = (torch.from_numpy(x) for x in (cnt, sm, mn, -mx))
    dist.all_reduce(tc, op=dist.ReduceOp.SUM); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
    dist.all_reduce(tmn, op=dist.ReduceOp.MIN); dist.all_reduce(tmx, op=dist.ReduceOp.MIN)   # MAX as MIN of the negation
    if rank == 0:
        out_q.put((gd[0].tolist(), gd[1].tolist(), tc.numpy().tolist(), ts.numpy().tolist(), tmn.numpy().tolist(), (-tmx.numpy()).tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_global_key_space_and_reduce():
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out_q)) for r in range(2)]
    for p in procs:
        p.start()
    d0, d1, cnt, sm, mn, mx = out_q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    q = parse_sql(SQL)
    segs = [datagen.make_segment_synth(i, 20_000, columns=COLS, vary_dim_dictionaries=True) for i in range(4)]
    exp = oracle.combine([oracle.execute(s, q) for s in segs])
    got = {}
    for slot, c in enumerate(cnt):
        if c:
            got[(d0[slot % len(d0)], d1[slot // len(d0)])] = [sm[slot], c, mn[slot], mx[slot]]
    assert set(got) == set(exp)
    for k, row in exp.items():
        assert got[k] == row, (k, got[k], row)
