"""GPU, >= 2 devices: the gathered pixels of a ray-sharded render (fused peer-store gather and NCCL all_gather) equal the 1-GPU render
bit for bit on every rank (BASELINE configs[3]).  Skipped on a single-GPU box; run with `gpurun --gpus 2`."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_render_gather_equals_single_gpu_render():
    n = min(torch.cuda.device_count(), 8)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                          "--master-port", "29611", str(ROOT / "tools" / "peer_gather_check.py")], capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "peer_gather_check ok" in out.stdout
