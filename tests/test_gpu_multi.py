"""Multi-GPU path on real devices (needs >= 2 GPUs in the box; skipped otherwise): the peer-memory halo
ring must reproduce the all_gather exchange bit for bit.  Host-side logic of the exchange is covered on
CPU by tests/test_dist_cpu.py (gloo)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_memory_halo_ring_matches_all_gather():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "mg_p2p_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert "p2p halo check: PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
