"""GPU parity tests for the remaining sparse formats of solve()'s accepted set: ell_matrix and hyb_matrix.
Everything goes through the C-ABI and is compared with golden vectors produced by the unmodified reference
(tests/golden/formats_vectors.npz, tests/golden/make_golden_formats.py)."""
import os

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu
MATS = ["lap2d_13x11", "cd3d_9x8x7", "ragged_200x180", "ragged_97x97"]


@pytest.fixture(scope="module")
def gf():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "formats_vectors.npz"))


def load_csr(golden, name):
    rows, cols = golden[name + "/shape"]
    return ol.CSR(rows, cols, golden[name + "/rp"], golden[name + "/ci"], golden[name + "/v"])


@pytest.mark.parametrize("name", MATS)
def test_ell_device_conversion_and_spmv_bitexact(pkg, be, golden, gf, name):
    A = load_csr(golden, name)
    dA = pkg.CsrMatrix.from_host(be, A.rows, A.cols, A.rp, A.ci, A.v)
    E = pkg.EllMatrix.from_csr(dA)                        # device-side conversion: the reference's layout exactly
    w, ir = gf[name + "/ell/width"]
    assert (E.width, E.internal_rows) == (w, ir)
    assert np.array_equal(E.coords.download()[:w * ir], gf[name + "/ell/coords"])
    assert np.array_equal(E.elements.download()[:w * ir], gf[name + "/ell/elements"])
    x, y0 = golden[name + "/x"], golden[name + "/y0"]
    dx = be.array(x)
    dy = be.array(np.full(A.rows, np.nan))                # beta == 0 must not read y
    E.spmv(dx, dy)
    assert np.array_equal(dy.download(), gf[name + "/ell/y"])
    dy = be.array(y0)
    E.spmv(dx, dy, 1.5, -0.25)
    assert ol.rel_err(dy.download(), gf[name + "/ell/y_ab"]).max() <= 1e-14
    # strided views: x at 3 + 2*col, y at 1 + 3*row (sparse.cpp:163-200); values equal the CSR golden result to rounding
    dxs, dys = be.array(golden[name + "/xs"]), be.array(golden[name + "/ys0"])
    E.spmv(dxs, dys, 1.0, 0.0, offx=3, incx=2, offy=1, incy=3)
    assert ol.rel_err(dys.download(), golden[name + "/ys"]).max() <= 1e-12


@pytest.mark.parametrize("name", MATS)
def test_hyb_device_conversion_and_spmv_bitexact(pkg, be, golden, gf, name):
    A = load_csr(golden, name)
    dA = pkg.CsrMatrix.from_host(be, A.rows, A.cols, A.rp, A.ci, A.v)
    H = pkg.HybMatrix.from_csr(dA)                        # csr_threshold 0.8 (hyb_matrix.hpp:44)
    w, ir, tn = gf[name + "/hyb/width"]
    assert (H.ell.width, H.ell.internal_rows, H.csr_nnz) == (w, ir, tn)
    assert np.array_equal(H.ell.coords.download()[:w * ir], gf[name + "/hyb/ell_coords"])
    assert np.array_equal(H.ell.elements.download()[:w * ir], gf[name + "/hyb/ell_elements"])
    assert np.array_equal(H.csr_rows.download(), gf[name + "/hyb/csr_rows"])
    assert np.array_equal(H.csr_cols.download()[:tn], gf[name + "/hyb/csr_cols"])
    assert np.array_equal(H.csr_elements.download()[:tn], gf[name + "/hyb/csr_elements"])
    x, y0 = golden[name + "/x"], golden[name + "/y0"]
    dx = be.array(x)
    dy = be.array(np.full(A.rows, np.nan))
    H.spmv(dx, dy)
    assert np.array_equal(dy.download(), gf[name + "/hyb/y"])
    dy = be.array(y0)
    H.spmv(dx, dy, 1.5, -0.25)
    assert ol.rel_err(dy.download(), gf[name + "/hyb/y_ab"]).max() <= 1e-14


def test_ell_padding_never_touches_x(pkg, be, orc):
    """Zero-valued slots must not propagate NaN/Inf from x[0] (cuda/sparse_matrix_operations.hpp:1777, host :1523)."""
    A = orc.stencil2d(9, 7)
    dA = pkg.CsrMatrix.from_host(be, A.rows, A.cols, A.rp, A.ci, A.v)
    x = orc.uniform(A.cols, 5, 1.0, 2.0)
    y_ref = orc.csr_spmv(A, x)
    xn = x.copy(); xn[0] = np.nan
    for M in (pkg.EllMatrix.from_csr(dA), pkg.HybMatrix.from_csr(dA)):
        dy = be.zeros(A.rows)
        M.spmv(be.array(xn), dy)
        y = dy.download()
        touched = np.isin(np.arange(A.rows), [0, 1, 9])   # rows that really reference column 0
        assert np.isnan(y[touched]).all() and not np.isnan(y[~touched]).any()
        assert ol.rel_err(y[~touched], y_ref[~touched]).max() <= 1e-13


@pytest.mark.parametrize("fmt", ["ell", "hyb"])
def test_solvers_on_ell_hyb_vs_reference(pkg, be, orc, golden, gf, fmt):
    """solve() on ell_matrix / hyb_matrix (cg.hpp:204-254 overloads): CG iteration counts within +-2 of the reference, same
    solution.  BiCGStab is chaotic: the reference ITSELF needs 77 (ELL), 86 (HYB) and a third count (CSR)
    iterations on the same convection-diffusion system, so the count must lie within +-4 of that spread."""
    for name, A, solvers in (("lap2d_63x65", orc.stencil2d(63, 65), ("cg", "bicgstab")), ("cd2d_48x50", orc.stencil2d(48, 50, 0.5, 0.0), ("bicgstab",))):
        b = np.ones(A.rows)
        dA = pkg.CsrMatrix.from_host(be, A.rows, A.cols, A.rp, A.ci, A.v)
        M = pkg.EllMatrix.from_csr(dA) if fmt == "ell" else pkg.HybMatrix.from_csr(dA, 0.8)
        db = be.array(b)
        for solver in solvers:
            dx = be.array(np.full(A.rows, 7.0))
            tag = pkg.SolverTag(tol=1e-8, max_iterations=1000).solve(solver, M, db, dx)
            key = "solve/%s/%s_%s" % (name, solver, fmt)
            it = int(gf[key + "/iters"][0])
            if solver == "cg":
                assert abs(tag.iters - it) <= 2, (key, tag.iters, it)
            else:
                spread = [int(gf["solve/%s/bicgstab_%s/iters" % (name, f)][0]) for f in ("ell", "hyb")]
                spread.append(int(golden["solve/%s/bicgstab_none/iters" % name][0]))
                assert min(spread) - 4 <= tag.iters <= max(spread) + 4, (key, tag.iters, spread)
            xr = gf[key + "/x"]
            assert np.linalg.norm(dx.download() - xr) <= 1e-5 * np.linalg.norm(xr)
        dx = be.zeros(A.rows)
        tag = pkg.SolverTag(tol=1e-8, max_iterations=900, krylov_dim=30).solve("gmres", M, db, dx)
        assert tag.error < 1e-8 and np.linalg.norm(b - A.to_scipy() @ dx.download()) / np.linalg.norm(b) < 1e-7


def test_ell_hyb_full_size_vs_csr(pkg, be):
    """256^3: ELL / HYB products equal the CSR product to rounding (different fma pattern, same order)."""
    A = pkg.CsrMatrix.stencil(be, 128, 128, 128)
    n = A.rows
    x, y0, y1, y2 = be.empty(n), be.zeros(n), be.zeros(n), be.zeros(n)
    be.check(be.L.ViennaCLCUDADfill_uniform(be.h, n, x.ptr, 3, 0, 1.0, 2.0))
    A.spmv(x, y0)
    E = pkg.EllMatrix.from_csr(A)
    assert E.width == 7
    E.spmv(x, y1)
    H = pkg.HybMatrix.from_csr(A, 0.01)                   # ELL width 6: interior rows spill into the CSR tail
    assert H.ell.width < 7 and H.csr_nnz > 1
    H.spmv(x, y2)
    ref = y0.download()
    # 6*x_i - (neighbours) cancels: compare against the scale of the terms (sum |a||x| <= 24), not of the result
    assert np.abs(y1.download() - ref).max() <= 24 * 1e-15 and np.abs(y2.download() - ref).max() <= 24 * 1e-15


@pytest.mark.parametrize("name", MATS)
def test_coo_index_and_spmv_bitexact(pkg, be, golden, gf, name):
    """coordinate_matrix: the CSR index built on the device equals the original CSR structure; the product is bit-identical
    to the reference's COO arithmetic for both the plain and the alpha/beta form."""
    A = load_csr(golden, name)
    M = pkg.CooMatrix(be, A.rows, A.cols, gf[name + "/coo/coords"], gf[name + "/coo/elements"])
    assert np.array_equal(M.index.rp.download(), A.rp)
    assert np.array_equal(M.index.ci.download()[:A.nnz], A.ci)
    x, y0 = golden[name + "/x"], golden[name + "/y0"]
    dx = be.array(x)
    dy = be.array(np.full(A.rows, np.nan))
    M.spmv(dx, dy)
    assert np.array_equal(dy.download(), gf[name + "/coo/y"])
    dy = be.array(y0)
    M.spmv(dx, dy, 1.5, -0.25)
    assert np.array_equal(dy.download(), gf[name + "/coo/y_ab"])
    # solvers run on the index like on any compressed_matrix
    if A.rows == A.cols and name.startswith("lap"):
        b = np.ones(A.rows)
        dsol = be.zeros(A.rows)
        tag = pkg.SolverTag(tol=1e-9, max_iterations=500).solve("cg", M.index, be.array(b), dsol)
        assert tag.error < 1e-9 and np.linalg.norm(b - A.to_scipy() @ dsol.download()) / np.linalg.norm(b) < 1e-8


def test_coo_unsorted_is_refused(pkg, be):
    coords = np.array([1, 0, 0, 1], np.uint32)           # row 1 before row 0
    with pytest.raises(pkg.VclError):
        pkg.CooMatrix(be, 2, 2, coords, np.array([1.0, 2.0]))


@pytest.mark.parametrize("scale", [0.5, 5.0])
def test_cg_jacobi_fused_vs_reference_generic_pcg(pkg, be, orc, gf, scale):
    """solve(A, b, cg_tag, jacobi_precond): fused single-reduction PCG on the device vs the reference's generic PCG
    (cg.hpp:257-322): iteration count within +-2, same solution, same error estimate definition."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mgf", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden_formats.py"))
    mgf = importlib.util.module_from_spec(spec); spec.loader.exec_module(mgf)
    A = mgf.vardiag_matrix(orc, scale)
    b = np.ones(A.rows)
    dA = pkg.CsrMatrix.from_host(be, A.rows, A.cols, A.rp, A.ci, A.v)
    dx = be.array(np.full(A.rows, 3.0))
    tag = pkg.SolverTag(tol=1e-9, max_iterations=2000, precond=1).solve("cg", dA, be.array(b), dx)
    key = "solve/vd2d_63x65_s%g/cg_jacobi" % scale
    it = int(gf[key + "/iters"][0])
    assert abs(tag.iters - it) <= 2, (tag.iters, it)
    xr = gf[key + "/x"]
    x = dx.download()
    assert np.linalg.norm(x - xr) <= 1e-6 * np.linalg.norm(xr)
    assert tag.error < 1e-9 and np.linalg.norm(b - A.to_scipy() @ x) / np.linalg.norm(b) < 1e-7
    # budget exhaustion reports the estimate and the iterate after exactly max_iterations updates
    tag = pkg.SolverTag(tol=1e-30, max_iterations=5, precond=1).solve("cg", dA, be.array(b), dx)
    assert tag.iters == 5 and tag.error > 1e-9


# ----------------------------------------------------------------------------------------------- SELL-C-sigma
def _sell_sigma_layout(A, Cs, sigma):
    """numpy restatement of ViennaCLCUDAcsr2sell_sigma: stable sort by decreasing row length inside windows of sigma rows, then
    the sigma = 1 slice layout (sliced_ell_matrix.hpp:140-214) over the permuted rows."""
    lens = np.diff(A.rp.astype(np.int64))
    ns = (A.rows - 1) // Cs + 1
    perm = np.full(ns * Cs, 0xFFFFFFFF, np.uint32)
    for w0 in range(0, A.rows, sigma):
        w1 = min(w0 + sigma, A.rows)
        order = np.argsort(-lens[w0:w1], kind="stable")
        perm[w0:w1] = (w0 + order).astype(np.uint32)
    cpb = np.zeros(ns, np.uint32); bs = np.zeros(ns, np.uint32)
    off = 0
    for s in range(ns):
        rows = perm[s * Cs:(s + 1) * Cs]
        rows = rows[rows != 0xFFFFFFFF]
        cpb[s] = lens[rows].max() if rows.size else 0
        bs[s] = off
        off += int(cpb[s]) * Cs
    ci = np.zeros(off, np.uint32); va = np.zeros(off, A.v.dtype)
    for i, r in enumerate(perm):
        if r == 0xFFFFFFFF:
            continue
        s, rib = divmod(i, Cs)
        k0, k1 = int(A.rp[r]), int(A.rp[r + 1])
        idx = int(bs[s]) + rib + Cs * np.arange(k1 - k0)
        ci[idx] = A.ci[k0:k1]; va[idx] = A.v[k0:k1]
    return perm, cpb, bs, ci, va, off


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("Cs,sigma", [(32, 256), (32, 4096), (8, 64), (64, 64)])
def test_sell_c_sigma_layout_and_product(pkg, be, dtype, Cs, sigma):
    """SELL-C-sigma: device-side window sort + conversion equal the numpy restatement array for array; the product
    (and the alpha/beta form) is bit-identical to the sigma = 1 product of the oracle -- the per-row arithmetic is the same,
    only the storage row differs; padding shrinks on an irregular matrix."""
    rng = np.random.default_rng(9)
    rows, cols = 5003, 4000
    lens = np.minimum((rng.pareto(1.3, rows) * 3).astype(np.int64), 300); lens[::11] = 0
    rp = np.zeros(rows + 1, np.uint32); rp[1:] = np.cumsum(lens)
    ci = np.concatenate([np.sort(rng.choice(cols, l, replace=False)) for l in lens]).astype(np.uint32)
    A = ol.CSR(rows, cols, rp, ci, rng.uniform(-1, 1, ci.size).astype(dtype), dtype)
    o = ol.oracle(dtype); o.set_threads(1)
    dA = pkg.CsrMatrix.from_host(be, A.rows, A.cols, A.rp, A.ci, A.v, dtype=dtype)
    S1 = dA.to_sell(Cs)
    S = dA.to_sell_sigma(Cs, sigma)
    perm, cpb, bs, sci, sva, tot = _sell_sigma_layout(A, Cs, sigma)
    assert S.padded_nnz == tot and (sigma == Cs or S.padded_nnz < 0.8 * S1.padded_nnz)    # sigma > C: most of the padding is gone
    assert np.array_equal(S.perm.download(), perm)
    assert np.array_equal(S.cpb.download()[:cpb.size], cpb) and np.array_equal(S.bs.download()[:bs.size], bs)
    assert np.array_equal(S.ci.download()[:tot], sci) and np.array_equal(S.va.download()[:tot], sva)
    x = o.uniform(cols, 5, 1.0, 2.0); y0 = o.uniform(rows, 6, -1.0, 1.0)
    ref1 = o.sell_spmv(o.sell_build(A, Cs), x)
    ref2 = o.sell_spmv(o.sell_build(A, Cs), x, y0.copy(), 1.5, -0.25)
    dx = be.array(x)
    dy = be.array(np.full(rows, np.nan, dtype))
    S1.spmv(dx, dy)                                      # sigma = 1 and sigma > 1 kernels in ONE process: they share a function-pointer
    assert np.array_equal(dy.download(), ref1)           # type, and each needs its own opt-in to > 48 KB of dynamic shared memory
    dy = be.array(np.full(rows, np.nan, dtype))
    S.spmv(dx, dy)
    assert np.array_equal(dy.download(), ref1)
    dy = be.array(y0)
    S.spmv(dx, dy, 1.5, -0.25)
    assert np.array_equal(dy.download(), ref2)


def test_sell_c_sigma_in_the_solvers(pkg, be, orc):
    """CG / BiCGStab / GMRES on a SELL-C-sigma matrix: same iteration counts as on the sigma = 1 matrix (identical products)."""
    A = orc.stencil2d(63, 65, 0.5, 0.0)
    L = orc.stencil2d(63, 65)
    b = np.ones(A.rows)
    for mat, solver, kw in ((L, "cg", {}), (A, "bicgstab", {}), (A, "gmres", dict(krylov_dim=30))):
        dA = pkg.CsrMatrix.from_host(be, mat.rows, mat.cols, mat.rp, mat.ci, mat.v)
        db, x1, x2 = be.array(b), be.zeros(mat.rows), be.zeros(mat.rows)
        t1 = pkg.SolverTag(tol=1e-8, max_iterations=900, **kw).solve(solver, dA.to_sell(32), db, x1)
        t2 = pkg.SolverTag(tol=1e-8, max_iterations=900, **kw).solve(solver, dA.to_sell_sigma(32, 512), db, x2)
        # identical products; the per-thread partial sums of the fused inner products are grouped differently (storage order)
        assert abs(t1.iters - t2.iters) <= 1 and np.allclose(x1.download(), x2.download(), rtol=1e-6, atol=1e-9), (solver, t1.iters, t2.iters)
