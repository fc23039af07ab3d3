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


def test_two_devices_one_process():
    """Two handles on two devices in ONE process (the occupancy / shared-memory opt-in cache was per process, so the second
    device's TMA kernels failed to launch; entry points now make the handle's device current)."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    pkg = ge.load_package()
    outs = []
    bes = [pkg.Backend(0), pkg.Backend(1)]
    mats = [pkg.CsrMatrix.stencil(b, 96, 96, 96) for b in bes]
    for rep in range(2):                                   # interleaved use of the two handles
        for b, A in zip(bes, mats):
            n = A.rows
            x = b.array(np.linspace(1.0, 2.0, n)); y = b.zeros(n)
            A.spmv(x, y)
            S = A.to_sell(32); ys = b.zeros(n)
            S.spmv(x, ys)
            rhs, sol = b.array(np.ones(n)), b.zeros(n)
            tag = pkg.SolverTag(tol=1e-8, max_iterations=500).solve("cg", A, rhs, sol)
            outs.append((y.download(), ys.download(), tag.iters, sol.download()))
    y0, ys0, it0, s0 = outs[0]
    for y, ys, it, s in outs[1:]:
        assert np.array_equal(y, y0) and np.array_equal(ys, ys0) and it == it0 and np.array_equal(s, s0)
    for b in bes:
        b.close()
