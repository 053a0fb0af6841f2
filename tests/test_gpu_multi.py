"""N > 1 parity (-m gpu, needs >= 2 GPUs; skipped on a single-GPU box): the merged table of a multi-GPU query must equal
the oracle's merge over all ranks' segments -- dense tables, DISTINCTCOUNT bitsets, filtered aggregations, keyless,
hash tables -- both deployments of include/pinot_b200.h: one process per GPU (NCCL inside libpinot_b200.so,
PB_Q_ALL_RANKS) and one process driving several GPUs (NVLink peer merge)."""
import os
import subprocess
import sys
import tempfile

import pytest

from pinot_b200 import native

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "multi_gpu_worker.py")


def _gpus():
    return native.lib().pb_device_count()


def _run(procs, timeout=600):
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for k in procs:
                k.kill()
            raise
        outs.append(out.decode("utf-8", "replace"))
    for p, out in zip(procs, outs):
        assert p.returncode == 0 and "MULTI_GPU_OK" in out, out[-4000:]


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 4])
def test_ranks_merge_inside_the_library_matches_oracle(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    with tempfile.TemporaryDirectory() as xdir:
        procs = [subprocess.Popen([sys.executable, WORKER, "ranks", str(r), str(world), xdir], cwd=ROOT,
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(world)]
        _run(procs)


@pytest.mark.timeout(900)
def test_one_process_driving_two_devices_matches_oracle():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    _run([subprocess.Popen([sys.executable, WORKER, "devices", "2"], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)])


@pytest.mark.timeout(600)
def test_single_rank_communicator_on_one_gpu():
    _run([subprocess.Popen([sys.executable, WORKER, "single"], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)])
