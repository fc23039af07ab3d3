"""CPU tests, single precision: the float build of the plain-C oracle (oracle/libvcl_oracle_f32.so) against the golden
vectors produced by the unmodified reference instantiated for NumericT = float (tests/golden/make_golden_f32.py), and --
where oracle/_ref was built -- against that reference itself on ragged matrices.  Every SpMV form is pinned BIT FOR BIT."""
import os

import numpy as np
import pytest

import oracle_lib as ol

F = np.float32
MATS = ["lap2d_13x11", "cd3d_9x8x7", "ragged_200x180", "ragged_97x97"]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def g32():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors_f32.npz"))


@pytest.fixture(scope="module")
def o32():
    o = ol.oracle(F)
    o.set_threads(1)
    return o


def load_csr(g, name):
    rows, cols = g[name + "/shape"]
    return ol.CSR(rows, cols, g[name + "/rp"], g[name + "/ci"], g[name + "/v"], F)


@pytest.mark.parametrize("name", MATS)
def test_csr_forms_bitexact_f32(g32, o32, name):
    A = load_csr(g32, name)
    x, y0 = g32[name + "/x"], g32[name + "/y0"]
    assert x.dtype == F and A.v.dtype == F
    assert np.array_equal(o32.csr_spmv(A, x, y0.copy()), g32[name + "/y_assign"])
    assert np.array_equal(o32.csr_spmv(A, x, y0.copy(), alpha=1.0, beta=1.0), g32[name + "/y_add"])
    assert np.array_equal(o32.csr_spmv(A, x, y0.copy(), alpha=-1.0, beta=1.0), g32[name + "/y_sub"])
    assert np.array_equal(o32.csr_spmv(A, x, y0.copy(), alpha=1.5, beta=-0.25), g32[name + "/y_ab"])
    ys = o32.csr_spmv(A, g32[name + "/xs"], g32[name + "/ys0"].copy(), offx=3, incx=2, offy=1, incy=3)
    assert np.array_equal(ys, g32[name + "/ys"])
    assert np.array_equal(o32.csr_diag(A), g32[name + "/diag"])


@pytest.mark.parametrize("name", MATS)
def test_other_formats_bitexact_f32(g32, o32, name):
    A = load_csr(g32, name)
    x, y0 = g32[name + "/x"], g32[name + "/y0"]
    S = o32.sell_build(A, 32)
    for k in ("cols_per_block", "block_start", "col_idx", "elements"):
        assert np.array_equal(S[k], g32[name + "/sell32/" + k]), k
    if name + "/sell32/y" in g32.files:
        assert np.array_equal(o32.sell_spmv(S, x), g32[name + "/sell32/y"])
        assert np.array_equal(o32.sell_spmv(S, x, y0.copy(), 1.5, -0.25), g32[name + "/sell32/y_ab"])
    E = o32.ell_build(A); H = o32.hyb_build(A); M = o32.coo_build(A)
    for fmt, fn, obj in (("ell", o32.ell_spmv, E), ("hyb", o32.hyb_spmv, H), ("coo", o32.coo_spmv, M)):
        assert np.array_equal(fn(obj, x), g32[name + "/" + fmt + "/y"]), fmt
        assert np.array_equal(fn(obj, x, y0.copy(), 1.5, -0.25), g32[name + "/" + fmt + "/y_ab"]), fmt


def test_blas1_f32(g32, o32):
    a, c = g32["blas1/a"], g32["blas1/c"]
    assert abs(o32.norm2(a) - g32["blas1/norm2"][0]) <= 1e-6 * g32["blas1/norm2"][0]
    assert abs(o32.inner_prod(a, c) - g32["blas1/inner"][0]) <= 1e-3      # plain float sums of 10007 terms


@pytest.mark.parametrize("name,solver,key", [("lap2d_63x65", "cg", "cg_none"), ("lap2d_63x65", "bicgstab", "bicgstab_none"),
                                             ("cd2d_48x50", "bicgstab", "bicgstab_none")])
def test_solver_counts_f32(g32, o32, name, solver, key):
    """Iteration counts of the float restatement vs the float reference (1 thread): CG +-2; BiCGStab within its own spread."""
    A = o32.stencil2d(63, 65) if name.startswith("lap") else o32.stencil2d(48, 50, 0.5, 0.0)
    b = np.ones(A.rows, F)
    res = getattr(o32, solver)(A, b, tol=1e-5, maxit=1000)
    want = int(g32["solve/%s/%s/iters" % (name, key)][0])
    assert abs(res["iters"] - want) <= (2 if solver == "cg" else 6), (res["iters"], want)
    xr = g32["solve/%s/%s/x" % (name, key)]
    assert np.linalg.norm(res["x"] - xr) <= 1e-3 * np.linalg.norm(xr)


@pytest.mark.skipif(not ol.have_ref(F), reason="oracle/_ref/libvcl_ref_f32.so not built")
def test_csr_live_reference_ragged_f32(o32):
    """Rows of 0..40 entries, every view form, 1 and 4 threads: bit-identical to the live float reference."""
    r = ol.ref(dtype=F)
    rng = np.random.default_rng(3)
    rows, cols = 3000, 2500
    lens = rng.integers(0, 41, rows); lens[::17] = 0
    rp = np.zeros(rows + 1, np.uint32); rp[1:] = np.cumsum(lens)
    ci = np.concatenate([np.sort(rng.choice(cols, l, replace=False)) for l in lens]).astype(np.uint32)
    A = ol.CSR(rows, cols, rp, ci, rng.uniform(-2, 2, ci.size).astype(F), F)
    for th in (1, 4):
        r.set_threads(th)
        for offx, incx, offy, incy in [(0, 1, 0, 1), (3, 2, 1, 3), (5, 1, 0, 1), (0, 1, 2, 2)]:
            x = o32.uniform(offx + cols * incx, 7, 1.0, 2.0)
            for alpha, beta in [(1, 0), (-1.25, 0.75)]:
                y0 = o32.uniform(offy + rows * incy, 9, -1, 1)
                yo = o32.csr_spmv(A, x, y0.copy(), alpha, beta, offx, incx, offy, incy)
                yr = r.csr_spmv(A, x, y0.copy(), alpha, beta, offx, incx, offy, incy, nx=cols, ny=rows)
                assert np.array_equal(yo, yr), (th, offx, incx, alpha, beta)
    r.set_threads(1)
