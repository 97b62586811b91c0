"""Layer-sharded edit on real GPUs (solver row): tests/tools/sharded_erase_check.py under torchrun with two ranks — the
erase driver with projections dealt over ranks and one exchange of the edited weights must reproduce the single-GPU artifact bit
for bit on every rank.  Skipped on a box with fewer than two devices (the CPU twin is tests/test_sharding_gloo.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two CUDA devices")
def test_sharded_erase_equals_single_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "tools", "sharded_erase_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert "sharded == single on every rank: True" in r.stdout
