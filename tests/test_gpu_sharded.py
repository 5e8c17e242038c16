"""GPU: the sharded mats axis (include/gpub200/sharded.cuh, gputils_b200/csrc/multi.cu). The C++ program
tests/host_harness/sharded_test.cu (built by __graft_entry__.build() into build/tests/) runs every sharded operation next to
the single-GPU DTensor path on the same inputs and demands bit-identical results -- shards are independent, so sharding
must not change a single bit -- for ragged and empty shards, with the result all-gathered over NCCL and over peer copies.
On a one-GPU box the same program runs with one shard and with two shards placed on the same device."""
import subprocess

import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu
BIN = REPO / "build" / "tests" / "sharded_test"


def _run(devices: str, transport: str):
    if not BIN.exists():
        pytest.fail(f"{BIN} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    r = subprocess.run([str(BIN), devices, transport, "4000"], capture_output=True, text=True, timeout=600)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "ALL PASSED" in out and "FAIL " not in out, out[-3000:]
    return out


def _device_count():
    import torch
    return torch.cuda.device_count()


def test_one_shard_equals_single_gpu_path():
    out = _run("0", "auto")
    assert "shards=1" in out


def test_two_shards_on_one_device_over_peer_copies():
    out = _run("0,0", "auto")
    assert "transport=p2p" in out          # NCCL refuses a device listed twice: the copy path serves it


def test_three_ragged_shards_on_one_device():
    _run("0,0,0", "p2p")


def test_two_devices_nccl_allgather():
    if _device_count() < 2:
        pytest.skip("needs two GPUs")
    out = _run("0,1", "nccl")
    assert "transport=nccl" in out


def test_two_devices_peer_copies():
    if _device_count() < 2:
        pytest.skip("needs two GPUs")
    _run("0,1", "p2p")


def test_all_devices_of_the_box():
    n = _device_count()
    if n < 3:
        pytest.skip("covered by the two-device tests")
    _run(",".join(str(i) for i in range(n)), "auto")
