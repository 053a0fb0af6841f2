"""The cross-GPU merge path on one GPU (-m gpu): PB_Q_COMBINE | PB_Q_DEFER_FINALIZE, NCCL all-reduce of the device tables
enqueued on the call's stream (world size 1 here; bench.py --gpus N runs the same code at N > 1), then finalize."""
import faulthandler
import os
import socket

import pytest

from oracle import oracle
from pinot_b200 import datagen, native
from pinot_b200.query import parse_sql
from tests.parity import assert_rows_equal, combined_rows

pytestmark = pytest.mark.gpu
faulthandler.enable()


def test_deferred_reduce_then_finalize():
    import torch
    import torch.distributed as dist
    from pinot_b200.distributed import agree_global_dictionaries, all_gather_merge_tables, all_reduce_tables

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        native.init(0)
        segs = [datagen.make_segment_synth(i, 70_000, columns=datagen.CONFIG2_COLUMNS, vary_dim_dictionaries=(i == 1)) for i in range(3)]
        staged = [native.StagedSegment(x) for x in segs]
        group = native.SegmentGroup(staged)
        q = parse_sql(datagen.config2_sql(segs, 200))
        agree_global_dictionaries(group, q.group_by, [0, 0, 0], dist)
        exp = combined_rows(oracle.combine([oracle.execute(x, q) for x in segs]), q)
        for it in range(4):
            r = native.execute(group, q, native.PB_Q_COMBINE | native.PB_Q_DEFER_FINALIZE)
            if it % 2 == 0:
                all_reduce_tables(r, q, dist, torch)
            else:
                all_gather_merge_tables(r, dist, torch)
            r.finalize()
            assert_rows_equal(r.tables[0].rows(), exp, q, exact_float=True, what="deferred + all-reduce")
            r.free()
    finally:
        dist.destroy_process_group()
