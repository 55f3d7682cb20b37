"""GPU (>= 2 devices): slab-sharded operator and Euler solve against the oracle / one GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("halo", ["push", "nccl"])
def test_sharded_matches_single_gpu(halo):
    """halo='push': NVLink peer-memory halo push (csrc/halo_push.cu, the default);
    halo='nccl': the all-gather fallback."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29533" if halo == "push" else "29534",
           os.path.join(ROOT, "tools", "check_sharded.py")]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, PDE_HALO=halo))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "MISMATCH" not in res.stdout
