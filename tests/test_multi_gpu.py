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
