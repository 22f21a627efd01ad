"""N > 1 on real GPUs: torchrun + NCCL, the sharded mesh must equal the 1-GPU mesh bit for
bit.  Skipped on boxes with a single GPU (the gloo tests in test_host_logic.py cover the
protocol there)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_extract_mesh_is_bit_exact(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, have {torch.cuda.device_count()}")
    cmd = [
        sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
        "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tools", "check_sharded.py"), "96",
    ]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("bit-exact vs 1 GPU: True") == 8 and "False" not in out.stdout, out.stdout
