"""Multi-GPU parity of the within-evaluation sharded path (2D block-cyclic Sigma_y, NCCL panel broadcasts).

Runs only on a box with >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests -m gpu`); on a single-GPU box the
sharded driver is still covered on a 1 x 1 grid (tests/test_gpu_parity.py) and its schedule by the 2-rank gloo
emulation (tests/test_dist_gloo.py).  Tolerance: 1e-13 relative on -2logL against the single-GPU path
(same per-tile DMMA order, so the results are expected to be bit-identical).
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world,n,p", [(2, 6000, 3), (2, 9100, 2), (4, 7000, 3), (8, 12000, 5)])
def test_sharded_equals_single_gpu(world, n, p):
    if _ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, NCCL_DEBUG="WARN")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", str(29600 + world),
                          os.path.join(ROOT, "tools", "dist_check.py"), str(n), str(p), "1"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "DIST_CHECK OK" in out.stdout, out.stdout[-2000:]
