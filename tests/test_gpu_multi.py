"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): tile-row bands reproduce the single-GPU / oracle
frame bit for bit, both with the fused peer-memory compositor store and with the NCCL gather."""
import os
import subprocess
import sys

import pytest

from godotgaussiansplatting_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_two_gpu_band_sharding_matches_oracle():
    ngpu = _lib.lib().gsr_device_count()
    if ngpu < 2:
        pytest.skip(f"needs 2 GPUs, box has {ngpu}")
    world = 4 if ngpu >= 4 else 2
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                          "--master-port", "29541", os.path.join(ROOT, "tests", "multi_gpu_worker.py")],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-5000:]
    assert "PEER_MODE_OK" in res.stdout and "NCCL_GATHER_OK" in res.stdout and "GROUP_MODE_OK" in res.stdout
