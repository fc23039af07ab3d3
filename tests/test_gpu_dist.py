"""Row-partitioned path.  The real check lives in tests/dist_check.py (one process per GPU, torchrun);
these wrappers run it with world = 1 always and with world = 2 when the box has two GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(cmd):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    sys.stdout.write(p.stdout[-4000:]); sys.stderr.write(p.stderr[-2000:])
    assert p.returncode == 0 and "DIST_CHECK PASS" in p.stdout


def test_dist_world1():
    _run([sys.executable, "tests/dist_check.py"])


def test_dist_world2():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run: gpurun --gpus 2 -- python -m pytest tests -m gpu -k dist)")
    _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
          "--master-port", "29517", "tests/dist_check.py"])
