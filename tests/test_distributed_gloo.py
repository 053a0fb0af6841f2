"""World-size-2 tests (CPU, no device) of the HOST side of the N > 1 path: the rendezvous helpers of
pinot_b200/distributed.py (torch.distributed `gloo` and the file exchange), the agreement on global dictionaries through
the C ABI (pb_segment_group_export_dictionary / _set_global_dictionary / _remap), and the static sharding of segments.
The device side of the merge (NCCL all-gather + pb_merge_blocks_kernel inside libpinot_b200.so) cannot run here; its
parity against the oracle at N = 2 and 4 is tests/test_gpu_multi.py (-m gpu) and bench.py's pre-timing check."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from pinot_b200 import datagen, native
from pinot_b200.distributed import (FileExchange, TorchExchange, agree_global_dictionaries, merge_sorted_dictionaries,
                                    shard_segments)

COLS = ["d0", "d1", "s0", "c3"]
TYPES = [0, 0, 4, 0]
N_SEGS = 5           # not a multiple of the world size: ranks hold different numbers of segments


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _segments():
    return [datagen.make_segment_synth(i, 4_000 + 500 * i, columns=COLS, vary_dim_dictionaries=True) for i in range(N_SEGS)]


def _check_rank(rank, world, exchange):
    segs = _segments()
    mine = [segs[i] for i in shard_segments(N_SEGS, rank, world)]
    group = native.SegmentGroup([native.StagedSegment(s) for s in mine])
    agree_global_dictionaries(group, COLS, TYPES, exchange)
    for col, ty in zip(COLS, TYPES):
        gd = group.export_dictionary(col)
        # the agreed dictionary is the sorted union over ALL ranks' segments ...
        if ty == 4:
            want = sorted({bytes(v) for s in segs for v in s.columns[col].dictionary_values()})
            got = [bytes(r).rstrip(b"\0") for r in gd]
            assert got == [w.rstrip(b"\0") for w in want], col
        else:
            want = np.unique(np.concatenate([s.columns[col].dictionary_values() for s in segs]))
            got = gd.reshape(-1).view(np.int32)
            assert np.array_equal(got, want), col
        # ... and every local dictId maps to the slot of its own value
        for si, s in enumerate(mine):
            rm = group.remap(col, si)
            local = s.columns[col].dictionary_values()
            assert len(rm) == len(local)
            if ty == 4:
                assert [got[g] for g in rm] == [bytes(v).rstrip(b"\0") for v in local], (col, si)
            else:
                assert np.array_equal(got[rm], local), (col, si)
    group.release()


def _gloo_worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _check_rank(rank, world, TorchExchange(dist))
        out_q.put((rank, "ok"))
    except Exception as e:      # pragma: no cover
        out_q.put((rank, repr(e)))
    dist.barrier()
    dist.destroy_process_group()


def _file_worker(rank, world, xdir, out_q):
    try:
        _check_rank(rank, world, FileExchange(xdir, rank, world))
        out_q.put((rank, "ok"))
    except Exception as e:      # pragma: no cover
        out_q.put((rank, repr(e)))


def _spawn(target, args_of):
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    procs = [ctx.Process(target=target, args=args_of(r) + (out_q,)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out_q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == {0: "ok", 1: "ok"}, got


@pytest.mark.timeout(300)
def test_two_ranks_agree_on_global_dictionaries_gloo():
    port = _free_port()
    _spawn(_gloo_worker, lambda r: (r, 2, port))


@pytest.mark.timeout(300)
def test_two_ranks_agree_on_global_dictionaries_file_exchange():
    with tempfile.TemporaryDirectory() as xdir:
        _spawn(_file_worker, lambda r: (r, 2, xdir))


def test_sharding_and_dictionary_merge():
    assert [shard_segments(5, r, 2) for r in range(2)] == [[0, 2, 4], [1, 3]]
    assert sorted(sum((shard_segments(64, r, 8) for r in range(8)), [])) == list(range(64))
    a = np.array([1, 5, 9], dtype=np.int32).view(np.uint8).reshape(-1, 4)
    b = np.array([2, 5], dtype=np.int32).view(np.uint8).reshape(-1, 4)
    assert merge_sorted_dictionaries([a, b], 0).reshape(-1).view(np.int32).tolist() == [1, 2, 5, 9]
    s1 = np.frombuffer(b"ab\0\0zz\0\0", dtype=np.uint8).reshape(2, 4)
    s2 = np.frombuffer(b"ab\0\0\0\0" + b"abc\0\0\0", dtype=np.uint8).reshape(2, 6)      # wider entries on another rank
    m = merge_sorted_dictionaries([s1, s2], 4)
    assert [bytes(r).rstrip(b"\0") for r in m] == [b"ab", b"abc", b"zz"] and m.shape[1] == 6


def test_comm_api_fails_cleanly_without_a_communicator():
    # no GPU here: the entry points must report, not crash (and PB_Q_ALL_RANKS is refused without pb_comm_init)
    has, n, r = native.comm_info()
    assert (has, n, r) == (False, 1, 0)
    with pytest.raises(native.PinotB200Error):
        native.comm_init(2, 5, b"\0" * native.PB_COMM_ID_BYTES)      # rank out of range
