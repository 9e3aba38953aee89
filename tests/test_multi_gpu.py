"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the peer-memory exchange
(pss_peer_search: publish / gather kernels over NVLink, one host synchronisation per query) and the host-driven
NCCL exchange against the single-GPU answer."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_and_nccl_exchange_match_single_gpu(world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "peer_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "mismatches=0" in p.stdout
    summary = [line for line in p.stdout.splitlines() if line.startswith("PEER_WORKER")]
    print("\n".join(summary))
    out_dir = os.path.join(ROOT, "gpurun_out")          # kept next to the other GPU-run artefacts (copied to profiles/)
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"peer_worker_{world}.log"), "w") as f:
            f.write(f"world={world}\n" + "\n".join(summary) + "\n")


def test_two_handles_on_two_devices_in_one_process():
    """kernel attributes (opt-in shared memory) and occupancy are per device; a second handle on another GPU of
    the same process must sieve, compact and search like the first"""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "prime-square-sum_b200"))
    from pss_b200 import _lib, search
    h0, h1 = _lib.Handle(0), _lib.Handle(1)
    try:
        assert "sm_10" in h0.device_info() and "sm_10" in h1.device_info()
        for h in (h1, h0, h1):
            assert h.count_primes(0, 10 ** 8) == 5761455
            assert int(h.generate_n_primes(100000)[-1]) == 1299709
            o = search.search(666, 10 ** 6, power=2, handle=h)
            assert o.first() == (7, 2) and o.last_prime == 15485863
        h0.set_search_path(True)
        assert search.search(8944, 10 ** 5, power=3, handle=h0).first() == (7, 3)
        assert search.search(8944, 10 ** 5, power=3, handle=h1).first() == (7, 3)
    finally:
        h0.close()
        h1.close()
