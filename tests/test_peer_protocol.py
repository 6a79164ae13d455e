"""The peer-memory exchange protocol of csrc/peer.cu (slots double-buffered by step parity,
release/acquire flags carrying the step number, the x_j area cleaned after the NEXT gather),
modelled with one host thread per rank (tests/sim/peer_protocol_sim.cpp) under random delays.
CPU only; complements the world_size-2 gloo test of the NCCL exchange.  The CUDA kernels
themselves are exercised by tests/test_gpu_multi.py (PRB_TEST_PEER=1) on >= 2 GPUs."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "sim", "peer_protocol_sim.cpp")


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path_factory.mktemp("peer") / "peer_sim")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, SRC], check=True)
    return exe


def _run(exe, *args):
    r = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=300)
    return r.returncode, r.stdout.strip()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_protocol_holds_under_random_delays(sim, world):
    for seed in (1, 2):
        rc, out = _run(sim, world, 384, 300, seed)
        assert rc == 0 and out == "ok", out


@pytest.mark.parametrize("mutation,what", [(1, "candidate slots not double-buffered"),
                                           (2, "x_j area cleaned before the step's gather")])
def test_model_detects_broken_variants(sim, mutation, what):
    """The model has teeth: each known-bad variant is caught for at least one seed."""
    for seed in (1, 2, 3):
        rc, out = _run(sim, 4, 384, 400, seed, mutation)
        if rc != 0 and "VIOLATION" in out:
            return
    pytest.fail("not detected: " + what)
