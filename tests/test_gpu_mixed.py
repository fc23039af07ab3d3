"""GPU parity: ViennaCLCUDADcsr_mixed_precision_cg against the reference's mixed_precision_cg.hpp (golden counts from the
reference host build, tests/golden/make_golden_mixed.py) and against the plain-C restatement oracle/vcl_oracle_mixed.c."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
CASES = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mixed_precision_cg.json")))


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%g-%d" % (c["name"], c["tol"], c["maxit"]))
def test_mixed_precision_cg_vs_reference(pkg, be, orc, case):
    """Same stopping rule and accounting as the reference: total (float) iterations, double error estimate = true residual.
    The inner iterations are float CG whose count depends on the summation order of its inner products (the reference's
    sequential float sums vs pairwise device sums, see tests/test_gpu_float.py), so the total is matched to ~10 %."""
    nx, ny, nz = case["grid"]
    A = orc.stencil3d(nx, ny, nz) if nz > 1 else orc.stencil2d(nx, ny)
    b = np.ones(A.rows)
    dA = pkg.CsrMatrix.from_host(be, A.rows, A.cols, A.rp, A.ci, A.v)
    db, dx = be.array(b), be.array(np.full(A.rows, 3.0))
    tag = pkg.mixed_precision_cg(dA, db, dx, case["tol"], case["maxit"], case["inner_tol"])
    x = dx.download()
    true = np.linalg.norm(b - A.to_scipy() @ x) / np.linalg.norm(b)
    if case["iters"] >= case["maxit"]:
        assert tag.iters == case["maxit"]                                 # budget exhausted: forced final update (:160)
        assert abs(tag.error - true) <= 1e-6 * true
        return
    assert tag.error < case["tol"] and abs(true - tag.error) <= 1e-3 * case["tol"] + 1e-12, (tag.error, true)
    assert abs(tag.iters - case["iters"]) <= max(3, 0.12 * case["iters"]), (tag.iters, case["iters"])
    assert abs(np.linalg.norm(x) - case["x_norm"]) <= 1e-7 * case["x_norm"]
    ref = orc.mixed_cg(A, b, case["tol"], case["maxit"], case["inner_tol"])
    assert np.linalg.norm(x - ref["x"]) <= 1e-7 * np.linalg.norm(ref["x"])


def test_mixed_precision_cg_with_kept_float_values_and_zero_rhs(pkg, be, orc):
    A = orc.stencil3d(20, 20, 20)
    dA = pkg.CsrMatrix.from_host(be, A.rows, A.cols, A.rp, A.ci, A.v)
    vals32 = be.array(A.v.astype(np.float32))                           # compressed_matrix<float> sharing A's index arrays
    b = orc.uniform(A.rows, 5, -1.0, 1.0)
    db, dx, dx2 = be.array(b), be.zeros(A.rows), be.zeros(A.rows)
    t1 = pkg.mixed_precision_cg(dA, db, dx, 1e-9, 1000, 1e-2)
    t2 = pkg.mixed_precision_cg(dA, db, dx2, 1e-9, 1000, 1e-2, values_float=vals32)
    assert t1.iters == t2.iters and np.array_equal(dx.download(), dx2.download())
    assert np.linalg.norm(b - A.to_scipy() @ dx.download()) / np.linalg.norm(b) < 1e-9
    dz = be.array(np.full(A.rows, 9.0))
    t3 = pkg.mixed_precision_cg(dA, be.zeros(A.rows), dz, 1e-9, 1000, 1e-2)     # zero rhs: zero solution (:114-115)
    assert t3.iters == 0 and not dz.download().any()


def test_convert(pkg, be):
    import ctypes as C
    x = np.linspace(-3, 3, 10001)
    dx, dy, dz = be.array(x), be.zeros(x.size, np.float32), be.zeros(x.size)
    be.check(be.L.ViennaCLCUDAconvert_DtoS(be.h, x.size, dx.ptr, dy.ptr))
    assert np.array_equal(dy.download(), x.astype(np.float32))
    be.check(be.L.ViennaCLCUDAconvert_StoD(be.h, x.size, dy.ptr, dz.ptr))
    assert np.array_equal(dz.download(), x.astype(np.float32).astype(np.float64))
