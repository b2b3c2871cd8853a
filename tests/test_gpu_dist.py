"""NCCL parity of the row-sharded search (needs >= 2 GPUs on the box; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("mode", ["fused", "peers", "reduce"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_nccl_bit_identical(world, mode):
    """fused: scan kernels post into rank 0's mailbox over NVLink (no collective), uneven shards;
    peers: NCCL all-gather + peer-memory MMR; reduce: NCCL all-gather + int32 reduce of pool rows."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + world + {"fused": 40, "peers": 0, "reduce": 20}[mode]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dist_parity.py"),
           "150001", "768"]
    env = dict(os.environ, RLR_DIST_MODE=mode)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert res.returncode == 0 and "DIST_PARITY_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
