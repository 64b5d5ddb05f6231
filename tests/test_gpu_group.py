"""Shard-group protocol (include/gsr.h gsr_group_*; csrc/group.cu) on ONE GPU: G contexts in one process, each on its own
stream, run the full NCCL-free multi-GPU frame path -- scatter projection with "peer" stores, flag words, rows composited into
rank 0's frames, pipelined read-back with slot release -- and must reproduce the oracle bit for bit.  The multi-process /
multi-GPU version of the same check is tests/test_gpu_multi.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("G,n,w,h,boost,overlap", [(2, 30000, 640, 360, 0.0, -1), (3, 30000, 800, 600, 0.0, -1), (8, 60000, 1280, 720, 1.2, -1), (5, 7, 100, 50, 0.0, -1),
                                                   (2, 30000, 640, 360, 0.0, 1), (3, 30000, 800, 600, 0.0, 1), (8, 60000, 1280, 720, 1.2, 1)])
def test_group_of_contexts_on_one_gpu_matches_oracle(G, n, w, h, boost, overlap):
    """overlap = 1: every rank's scatter projection of frame f+1 runs on its front stream beside the compositor of frame f (three arena phases)."""
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "group_inprocess_worker.py"), str(G), str(n), str(w), str(h), str(boost), str(overlap)],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert res.returncode == 0 and f"GROUP_INPROCESS_OK G={G}" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
