"""Multi-GPU slab decomposition against the single-GPU result and the oracle (needs >= 2 GPUs; run on
the box with `gpurun --gpus 2`).  Launches tests/slab_worker.py under torchrun."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("scene,steps", [("jelly", 30), ("jelly_shear", 30), ("jelly_rebalance", 30), ("split_layers", 25), ("sand", 40), ("jelly_adaptive", 40), ("energy_error", 6)])
def test_slabs_match_single_gpu(scene, steps):
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "slab_worker.py"), scene, str(steps)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-6000:] + r.stderr[-1500:]
    assert "within tolerance" in r.stdout
