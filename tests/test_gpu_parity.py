"""GPU parity tests (run with -m gpu on a B200).  Everything goes through the C-ABI (include/vcl_b200.h) and is compared with
the CPU oracle / the golden vectors produced by the reference.  Structure follows the reference's tests/src/sparse.cpp:
plain product, += / -= forms, strided/ranged vectors, other formats, plus the solver runs the reference never tested."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

MATS = ["lap2d_13x11", "cd3d_9x8x7", "ragged_200x180", "ragged_97x97"]
EPS = 1e-12   # tests/src/sparse.cpp:1113 (double)


def load_csr(golden, name):
    rows, cols = golden[name + "/shape"]
    return ol.CSR(rows, cols, golden[name + "/rp"], golden[name + "/ci"], golden[name + "/v"])


def dev_csr(pkg, be, A, with_blocks=True):
    return pkg.CsrMatrix.from_host(be, A.rows, A.cols, A.rp, A.ci, A.v, with_blocks=with_blocks)


# ----------------------------------------------------------------------------------------------- SpMV, golden vectors
@pytest.mark.parametrize("use_blocks", [True, False])
@pytest.mark.parametrize("name", MATS)
def test_csr_spmv_golden_bitexact(pkg, be, golden, name, use_blocks):
    A = load_csr(golden, name)
    dA = dev_csr(pkg, be, A)
    x, y0 = golden[name + "/x"], golden[name + "/y0"]
    dx = be.array(x)
    # rows of more than 64 entries (CSR_LONG_ROW) are summed by a whole warp on the row-block path: tolerance-level parity there
    exact = np.diff(A.rp.astype(np.int64)) <= (64 if use_blocks else 1 << 30)
    for key, (alpha, beta) in {"y_assign": (1.0, 0.0), "y_add": (1.0, 1.0), "y_sub": (-1.0, 1.0)}.items():
        dy = be.array(y0)
        dA.spmv(dx, dy, alpha, beta, use_blocks=use_blocks)
        y = dy.download()
        assert np.array_equal(y[exact], golden[name + "/" + key][exact]), key
        assert ol.rel_err(y, golden[name + "/" + key]).max() <= 1e-14, key
    dy = be.array(y0)
    dA.spmv(dx, dy, 1.5, -0.25, use_blocks=use_blocks)
    assert ol.rel_err(dy.download(), golden[name + "/y_ab"]).max() <= 1e-14
    # strided / ranged views (sparse.cpp:163-200, :397-400)
    dxs, dys = be.array(golden[name + "/xs"]), be.array(golden[name + "/ys0"])
    dA.spmv(dxs, dys, 1.0, 0.0, offx=3, incx=2, offy=1, incy=3, use_blocks=use_blocks)
    ys, ys_ref = dys.download(), golden[name + "/ys"]
    touched = np.zeros(ys.size, bool); touched[1 + 3 * np.nonzero(~exact)[0]] = True      # y entries of the long rows
    assert np.array_equal(ys[~touched], ys_ref[~touched]) and ol.rel_err(ys, ys_ref).max() <= 1e-14


def test_beta_zero_does_not_read_y(pkg, be, golden):
    """spmv_alpha_beta: beta == 0 must not propagate NaNs from an uninitialised y (cuda/sparse_matrix_operations.hpp:130)."""
    A = load_csr(golden, "lap2d_13x11")
    dA = dev_csr(pkg, be, A)
    dx = be.array(golden["lap2d_13x11/x"])
    dy = be.array(np.full(A.rows, np.nan))
    dA.spmv(dx, dy)
    assert np.array_equal(dy.download(), golden["lap2d_13x11/y_assign"])
    dS = dA.to_sell(32)
    dy = be.array(np.full(A.rows, np.nan))
    dS.spmv(dx, dy)
    assert not np.isnan(dy.download()).any()


def test_alias_is_refused_at_the_abi(pkg, be, golden):
    """x = A*x is resolved by the facade through a temporary (compressed_matrix.hpp:1237-1242); the raw ABI refuses it."""
    A = load_csr(golden, "ragged_97x97")
    dA = dev_csr(pkg, be, A)
    dx = be.array(golden["ragged_97x97/x"])
    with pytest.raises(pkg.VclError):
        dA.spmv(dx, dx)
    # the facade's recipe: temporary, then copy back
    tmp = be.zeros(A.rows)
    dA.spmv(dx, tmp)
    assert np.array_equal(tmp.download(), golden["ragged_97x97/x_alias"])


@pytest.mark.parametrize("name", MATS)
def test_sell_build_and_spmv_golden(pkg, be, golden, name):
    A = load_csr(golden, name)
    dA = dev_csr(pkg, be, A)
    dS = dA.to_sell(32)                                   # device-side conversion
    assert np.array_equal(dS.cpb.download()[:len(golden[name + "/sell32/cols_per_block"])], golden[name + "/sell32/cols_per_block"])
    assert np.array_equal(dS.bs.download()[:len(golden[name + "/sell32/block_start"])], golden[name + "/sell32/block_start"])
    n = dS.padded_nnz
    assert n == len(golden[name + "/sell32/elements"])
    assert np.array_equal(dS.ci.download()[:n], golden[name + "/sell32/col_idx"])
    assert np.array_equal(dS.va.download()[:n], golden[name + "/sell32/elements"])
    x, y0 = golden[name + "/x"], golden[name + "/y0"]
    dx, dy = be.array(x), be.array(y0)
    dS.spmv(dx, dy)
    y = dy.download()
    if name + "/sell32/y" in golden.files:
        assert np.array_equal(y, golden[name + "/sell32/y"])
        dy = be.array(y0)
        dS.spmv(dx, dy, 1.5, -0.25)
        assert ol.rel_err(dy.download(), golden[name + "/sell32/y_ab"]).max() <= 1e-14
    assert ol.rel_err(y, golden[name + "/y_assign"]).max() <= EPS


@pytest.mark.parametrize("name", MATS)
def test_row_info_diag_golden(pkg, be, golden, name):
    A = load_csr(golden, name)
    dA = dev_csr(pkg, be, A)
    assert np.array_equal(dA.row_info(3).download(), golden[name + "/diag"])
    M = A.to_scipy()
    assert np.allclose(dA.row_info(0).download(), abs(M).max(axis=1).toarray().ravel())
    assert np.allclose(dA.row_info(1).download(), np.asarray(abs(M).sum(axis=1)).ravel())
    assert np.allclose(dA.row_info(2).download(), np.sqrt(np.asarray(M.multiply(M).sum(axis=1)).ravel()))


# ----------------------------------------------------------------------------------------------- SpMV vs oracle, edge cases
def test_empty_and_tiny(pkg, be, orc):
    L = be.L
    assert L.ViennaCLCUDADcsrmv(be.h, 0, 0, 0, None, None, None, None, 0, None, 0, 1, 1.0, None, 0, 1, 0.0) == 0
    A = ol.CSR(3, 3, [0, 0, 0, 0], [], [])               # all rows empty
    dA = dev_csr(pkg, be, A)
    dx, dy = be.array(np.ones(3)), be.array(np.full(3, 7.0))
    dA.spmv(dx, dy)
    assert np.array_equal(dy.download(), np.zeros(3))
    A = ol.CSR(1, 1, [0, 1], [0], [2.5])
    dA = dev_csr(pkg, be, A)
    dx, dy = be.array(np.array([4.0])), be.zeros(1)
    dA.spmv(dx, dy)
    assert dy.download()[0] == 10.0


def test_long_rows_and_many_empty_rows(pkg, be, orc):
    """Rows longer than a row block (whole-CTA path), runs of empty rows, rectangular shape."""
    rng = np.random.default_rng(5)
    rows, cols = 3000, 9000
    rp = [0]; ci = []; v = []
    for r in range(rows):
        if r in (10, 11, 1500):
            n = 6000 if r != 11 else 2049
        elif 100 <= r < 700:
            n = 0
        else:
            n = int(rng.integers(0, 40))
        c = np.sort(rng.choice(cols, size=n, replace=False))
        ci.extend(c.tolist()); v.extend(rng.uniform(-1, 1, n).tolist()); rp.append(len(ci))
    A = ol.CSR(rows, cols, rp, ci, v)
    x = orc.uniform(cols, 17, 1.0, 2.0)
    y_ref = orc.csr_spmv(A, x)
    dA = dev_csr(pkg, be, A)
    dx, dy = be.array(x), be.zeros(rows)
    dA.spmv(dx, dy)
    y = dy.download()
    long_rows = np.array([10, 11, 1500])
    short = np.ones(rows, bool); short[long_rows] = False
    assert np.array_equal(y[short], y_ref[short])                       # sequential rows: bit-exact
    assert np.abs(y[long_rows] - y_ref[long_rows]).max() <= 1e-11       # tree-summed long rows: |terms| ~ 6000
    dy2 = be.zeros(rows)
    dA.spmv(dx, dy2, use_blocks=False)
    assert np.array_equal(dy2.download(), y_ref)
    dS = dA.to_sell(32)
    S = orc.sell_build(A, 32)
    dy3 = be.zeros(rows)
    dS.spmv(dx, dy3)
    assert np.array_equal(dy3.download(), orc.sell_spmv(S, x))


@pytest.mark.parametrize("shape", [(64, 64, 64), (100, 37, 29), (255, 255, 1), (1024, 64, 1)])
def test_stencil_generator_and_spmv_vs_oracle(pkg, be, orc, shape):
    nx, ny, nz = shape
    c = (0.5, 0.25, 0.125)
    A = orc.stencil3d(nx, ny, nz, *c) if nz > 1 else orc.stencil2d(nx, ny, c[0], c[1])
    dA = pkg.CsrMatrix.stencil(be, nx, ny, nz, *c)
    assert dA.nnz == A.nnz
    assert np.array_equal(dA.rp.download(), A.rp)
    assert np.array_equal(dA.ci.download()[:A.nnz], A.ci)
    assert np.array_equal(dA.va.download()[:A.nnz], A.v)
    x = orc.uniform(A.cols, 1, 1.0, 2.0)
    dx = be.empty(A.cols)
    be.check(be.L.ViennaCLCUDADfill_uniform(be.h, A.cols, dx.ptr, 1, 0, 1.0, 2.0))
    assert np.array_equal(dx.download(), x)
    orc.set_threads(orc.max_threads())
    y_ref = orc.csr_spmv(A, x)
    S = orc.sell_build(A, 32)
    ys_ref = orc.sell_spmv(S, x)
    orc.set_threads(1)
    dy = be.zeros(A.rows)
    dA.spmv(dx, dy)
    assert np.array_equal(dy.download(), y_ref)
    dS = dA.to_sell(32)
    assert dS.padded_nnz == S["padded_nnz"]
    dy.fill0()
    dS.spmv(dx, dy)
    ys = dy.download()
    assert np.array_equal(ys, ys_ref)
    # cross-format: the reference build fuses SELL's multiply-adds but not CSR's, so CSR and SELL differ by rounding of
    # O(eps * sum|a_ij x_j|) (~ 24 * 1.1e-16 here) even on the host; entries that cancel make this large in RELATIVE terms
    assert np.abs(ys - y_ref).max() <= 1e-13 * 24


@pytest.mark.parametrize("Cs", [4, 32, 48, 64, 256, 512])
def test_sell_other_slice_heights(pkg, be, orc, Cs):
    """rows_per_block other than the default 32 (sliced_ell_matrix.hpp:59-63), incl. the non-staged kernel paths."""
    A = orc.stencil3d(19, 17, 23, 0.5, 0.25, 0.125)
    x = orc.uniform(A.cols, 9, 1.0, 2.0)
    S = orc.sell_build(A, Cs)
    dA = dev_csr(pkg, be, A)
    dS = dA.to_sell(Cs)
    assert dS.padded_nnz == S["padded_nnz"]
    assert np.array_equal(dS.va.download()[:S["padded_nnz"]], S["elements"])
    dx, dy = be.array(x), be.array(np.full(A.rows, np.nan))
    dS.spmv(dx, dy)
    assert np.array_equal(dy.download(), orc.sell_spmv(S, x))
    dy = be.array(x[:A.rows].copy())
    dS.spmv(dx, dy, 0.5, 2.0)
    assert ol.rel_err(dy.download(), orc.sell_spmv(S, x, x[:A.rows].copy(), 0.5, 2.0)).max() <= 1e-14
    dS2 = pkg.SellMatrix.from_host(be, S)                 # arrays uploaded in the reference layout
    dy.fill0()
    dS2.spmv(dx, dy)
    assert np.array_equal(dy.download(), orc.sell_spmv(S, x))


def test_partial_row_generator(pkg, be, orc):
    nx, ny, nz = 20, 15, 12
    A = orc.stencil3d(nx, ny, nz, 0.1, 0.2, 0.3)
    rb, re_ = 1234, 2999
    dA = pkg.CsrMatrix.stencil(be, nx, ny, nz, 0.1, 0.2, 0.3, row_begin=rb, row_end=re_)
    rp = dA.rp.download()
    assert np.array_equal(rp, A.rp[rb:re_ + 1] - A.rp[rb])
    assert np.array_equal(dA.ci.download()[:dA.nnz], A.ci[A.rp[rb]:A.rp[re_]])
    assert np.array_equal(dA.va.download()[:dA.nnz], A.v[A.rp[rb]:A.rp[re_]])


def test_full_size_spmv_properties(pkg, be, orc):
    """BASELINE config 2 (3-D 7-point Laplacian 256^3, 16.7M rows, 117M nnz): size-independent properties.
    A*1 counts the missing (Dirichlet) neighbours exactly; CSR and SELL agree; linearity holds to rounding."""
    n1 = 256
    dA = pkg.CsrMatrix.stencil(be, n1, n1, n1)
    N = n1 ** 3
    assert dA.nnz == 7 * N - 6 * n1 * n1
    ones, dy = be.array(np.ones(N)), be.zeros(N)
    dA.spmv(ones, dy)
    y = dy.download().reshape(n1, n1, n1)
    idx = np.arange(n1)
    edge = ((idx == 0) | (idx == n1 - 1)).astype(np.float64)
    expect = edge[:, None, None] + edge[None, :, None] + edge[None, None, :]
    assert np.array_equal(y, expect)
    dS = dA.to_sell(32)
    assert dS.padded_nnz == 117178368                                   # SELL-32 padded nnz at 256^3
    dx = be.empty(N)
    be.check(be.L.ViennaCLCUDADfill_uniform(be.h, N, dx.ptr, 3, 0, 1.0, 2.0))
    dy2 = be.zeros(N)
    dA.spmv(dx, dy)
    dS.spmv(dx, dy2)
    y_csr, y_sell = dy.download(), dy2.download()
    assert np.abs(y_csr - y_sell).max() <= 1e-13 * 24                 # see test_stencil_generator_and_spmv_vs_oracle
    assert np.median(ol.rel_err(y_csr, y_sell)) <= 1e-15
    # spot-check 100k rows against the oracle arithmetic on the host (rows regenerated independently)
    x = dx.download().reshape(n1, n1, n1)
    rng = np.random.default_rng(0)
    l, j, i = (rng.integers(1, n1 - 1, 100000) for _ in range(3))
    acc = np.zeros(100000)
    for (dl, dj, di, val) in ((-1, 0, 0, -1.0), (0, -1, 0, -1.0), (0, 0, -1, -1.0), (0, 0, 0, 6.0), (0, 0, 1, -1.0), (0, 1, 0, -1.0), (1, 0, 0, -1.0)):
        acc = acc + val * x[l + dl, j + dj, i + di]                     # mul, then add: the oracle's CSR chain
    assert np.array_equal(y_csr.reshape(n1, n1, n1)[l, j, i], acc)


# ----------------------------------------------------------------------------------------------- BLAS-1
def test_blas1_golden(pkg, be, golden):
    a, c = golden["blas1/a"], golden["blas1/c"]
    da, dc = be.array(a), be.array(c)
    r = C.c_double(0)
    be.check(be.L.ViennaCLCUDADnrm2(be.h, a.size, C.byref(r), da.ptr, 0, 1))
    assert abs(r.value - golden["blas1/norm2"][0]) <= 1e-13 * golden["blas1/norm2"][0]
    be.check(be.L.ViennaCLCUDADdot(be.h, a.size, C.byref(r), da.ptr, 0, 1, dc.ptr, 0, 1))
    assert abs(r.value - golden["blas1/inner"][0]) <= 1e-12
    # strided dot: every 3rd element starting at 2
    n = (a.size - 2 + 2) // 3
    be.check(be.L.ViennaCLCUDADdot(be.h, n, C.byref(r), da.ptr, 2, 3, dc.ptr, 2, 3))
    assert abs(r.value - np.dot(a[2::3][:n], c[2::3][:n])) <= 1e-12
    dz = be.zeros(a.size)
    be.check(be.L.ViennaCLCUDADavbv(be.h, a.size, dz.ptr, 0, 1, da.ptr, 0, 1, 2.0, dc.ptr, 0, 1, -3.0))
    assert np.allclose(dz.download(), 2.0 * a - 3.0 * c, rtol=1e-15, atol=1e-15)
    be.check(be.L.ViennaCLCUDADavbv_v(be.h, a.size, dz.ptr, 0, 1, da.ptr, 0, 1, 1.0, dc.ptr, 0, 1, 1.0))
    assert np.allclose(dz.download(), 3.0 * a - 2.0 * c, rtol=1e-14, atol=1e-15)
    be.check(be.L.ViennaCLCUDADav(be.h, a.size, dz.ptr, 0, 1, da.ptr, 0, 1, 0.5))
    assert np.array_equal(dz.download(), 0.5 * a)
    be.check(be.L.ViennaCLCUDADelement_div(be.h, a.size, dz.ptr, 0, 1, da.ptr, 0, 1, dc.ptr, 0, 1))
    assert np.array_equal(dz.download(), a / c)
    be.check(be.L.ViennaCLCUDADassign(be.h, 10, dz.ptr, 5, 2, 9.0))
    z = dz.download()
    assert np.all(z[5:25:2] == 9.0) and z[6] != 9.0


# ----------------------------------------------------------------------------------------------- fused steps (per-op ABI)
def test_fused_steps_vs_numpy(pkg, be, orc):
    A = orc.stencil3d(23, 19, 17, 0.3, 0.2, 0.1)
    n = A.rows
    M = A.to_scipy()
    dA = dev_csr(pkg, be, A); dS = dA.to_sell(32)
    rng = np.random.default_rng(1)
    p, r0 = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    dp, dr0, dAp = be.array(p), be.array(r0), be.zeros(n)
    L = be.L
    for fmt in ("csr", "sell"):
        mat = dA if fmt == "csr" else dS
        s = mat.struct()
        buf = be.zeros(3 * 256)
        be.check(getattr(L, "ViennaCLCUDADpipelined_cg_prod_" + fmt)(be.h, C.byref(s), dp.ptr, dAp.ptr, buf.ptr, 768))
        Ap = dAp.download(); hb = buf.download()
        assert ol.rel_err(Ap, M @ p).max() <= 1e-11
        assert abs(hb[256:512].sum() - Ap @ Ap) <= 1e-10 * (Ap @ Ap)
        assert abs(hb[512:768].sum() - p @ Ap) <= 1e-10 * abs(p @ Ap)
        assert hb[0] == 0.0 and hb[257] == 0.0
        buf = be.zeros(6 * 256)
        be.check(getattr(L, "ViennaCLCUDADpipelined_bicgstab_prod_" + fmt)(be.h, C.byref(s), dp.ptr, dAp.ptr, dr0.ptr, buf.ptr, 256, 3 * 256))
        hb = buf.download()
        assert abs(hb[768:1024].sum() - Ap @ r0) <= 1e-10 * max(1.0, abs(Ap @ r0))
    # CG vector update
    x, r = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    Ap = M @ p
    dx, dr, dAp = be.array(x), be.array(r), be.array(Ap)
    buf = be.zeros(768)
    be.check(L.ViennaCLCUDADpipelined_cg_vector_update(be.h, n, dx.ptr, 0.7, dp.ptr, dr.ptr, dAp.ptr, -0.3, buf.ptr, 768))
    r_new = r - 0.7 * Ap
    assert np.allclose(dx.download(), x + 0.7 * p, rtol=1e-14, atol=1e-15)
    assert np.allclose(dr.download(), r_new, rtol=1e-14, atol=1e-15)
    assert np.allclose(dp.download(), r_new + (-0.3) * p, rtol=1e-14, atol=1e-15)
    assert abs(buf.download()[:256].sum() - r_new @ r_new) <= 1e-11 * (r_new @ r_new)
    # BiCGStab update_s: alpha = sum(chunk0) / sum(chunk3), partials spread over the chunks by a foreign producer
    hb = np.zeros(6 * 256); hb[0:256] = 0.5 / 256; hb[768:1024] = 2.0 / 256
    buf = be.array(hb)
    ds = be.zeros(n)
    dr = be.array(r)
    be.check(L.ViennaCLCUDADpipelined_bicgstab_update_s(be.h, n, ds.ptr, dr.ptr, dAp.ptr, buf.ptr, 256, 5 * 256))
    s_new = r - 0.25 * Ap
    assert np.allclose(ds.download(), s_new, rtol=1e-13, atol=1e-15)
    assert abs(buf.download()[5 * 256] - s_new @ s_new) <= 1e-11 * (s_new @ s_new)
    # BiCGStab vector update
    As = M @ s_new
    p0 = rng.uniform(-1, 1, n)
    dx, dp, dr, dAs = be.array(x), be.array(p0), be.array(r), be.array(As)
    buf = be.zeros(6 * 256)
    be.check(L.ViennaCLCUDADpipelined_bicgstab_vector_update(be.h, n, dx.ptr, 0.4, dp.ptr, 0.6, ds.ptr, dr.ptr, dAs.ptr, -0.2,
                                                              dAp.ptr, dr0.ptr, buf.ptr, 256))
    r2 = s_new - 0.6 * As
    assert np.allclose(dx.download(), x + 0.4 * p0 + 0.6 * s_new, rtol=1e-13, atol=1e-15)
    assert np.allclose(dr.download(), r2, rtol=1e-13, atol=1e-15)
    assert np.allclose(dp.download(), r2 + (-0.2) * (p0 - 0.6 * Ap), rtol=1e-13, atol=1e-14)
    assert abs(buf.download()[0] - r2 @ r0) <= 1e-11 * max(1.0, abs(r2 @ r0))


@pytest.mark.parametrize("m,k", [(12, 7), (64, 63)])
def test_gmres_steps_vs_numpy(pkg, be, orc, m, k):
    """(64, 63): the largest Krylov dimension -- stage 2 folds 63 chunk sums into device scratch (that scratch overflowed
    from k = 49 on); the state after it must still be intact, which the solves of the later tests on the same handle rely on."""
    n = 5003
    isz = (n + 127) // 128 * 128
    rng = np.random.default_rng(2)
    V = np.zeros(isz * m)
    for j in range(k + 1):
        V[j * isz:j * isz + n] = rng.uniform(-1, 1, n)
    res = rng.uniform(-1, 1, n)
    dV, dres = be.array(V), be.array(res)
    L = be.L
    chunk = 128
    dh = be.zeros(chunk * m)
    be.check(L.ViennaCLCUDADpipelined_gmres_gram_schmidt_stage1(be.h, dV.ptr, n, isz, k, dh.ptr, chunk))
    h = dh.download().reshape(m, chunk).sum(axis=1)
    vk = V[k * isz:k * isz + n]
    h_ref = np.array([V[j * isz:j * isz + n] @ vk for j in range(k)])
    assert np.allclose(h[:k], h_ref, rtol=1e-12, atol=1e-12)
    dR, dbuf = be.zeros(m * m), be.zeros(3 * chunk)
    be.check(L.ViennaCLCUDADpipelined_gmres_gram_schmidt_stage2(be.h, dV.ptr, n, isz, k, dh.ptr, dR.ptr, m, dbuf.ptr, chunk))
    vk2 = vk - sum(h_ref[j] * V[j * isz:j * isz + n] for j in range(k))
    Vd = dV.download()
    assert np.allclose(Vd[k * isz:k * isz + n], vk2, rtol=1e-12, atol=1e-13)
    R = dR.download()
    assert np.allclose(R[np.arange(k) + k * m], h_ref, rtol=1e-12, atol=1e-12)
    nsq = dbuf.download()[chunk:2 * chunk].sum()
    assert abs(nsq - vk2 @ vk2) <= 1e-11 * (vk2 @ vk2)
    dxi = be.zeros(chunk * m)
    be.check(L.ViennaCLCUDADpipelined_gmres_normalize_vk(be.h, n, dV.at(k * isz), dres.ptr, dR.ptr, k * m + k, dbuf.ptr, dxi.ptr, chunk, k * chunk))
    nrm = np.sqrt(vk2 @ vk2)
    Vd = dV.download()
    assert np.allclose(Vd[k * isz:k * isz + n], vk2 / nrm, rtol=1e-12, atol=1e-14)
    assert abs(dR.download()[k * m + k] - nrm) <= 1e-12 * nrm
    assert abs(dxi.download()[k * chunk] - res @ (vk2 / nrm)) <= 1e-11
    coef = rng.uniform(-1, 1, m)
    x = rng.uniform(-1, 1, n)
    dx, dc = be.array(x), be.array(coef)
    be.check(L.ViennaCLCUDADpipelined_gmres_update_result(be.h, n, dx.ptr, dres.ptr, dV.ptr, isz, dc.ptr, 5))
    x_ref = x + coef[0] * res + sum(coef[j] * Vd[(j - 1) * isz:(j - 1) * isz + n] for j in range(1, 5))
    assert np.allclose(dx.download(), x_ref, rtol=1e-13, atol=1e-14)


# ----------------------------------------------------------------------------------------------- solvers
def _mat(orc, name):
    if name == "lap2d_63x65":
        return orc.stencil2d(63, 65)
    if name == "cd2d_48x50":
        return orc.stencil2d(48, 50, 0.5, 0.0)
    return orc.stencil3d(11, 10, 9, 0.5, 0.25, 0.125)


def _true_res(A, b, x):
    return np.linalg.norm(b - A.to_scipy() @ x) / np.linalg.norm(b)


@pytest.mark.parametrize("fmt", ["csr", "sell"])
@pytest.mark.parametrize("name", ["lap2d_63x65", "cd2d_48x50", "cd3d_11x10x9"])
def test_solvers_vs_reference_golden(pkg, be, orc, golden, name, fmt):
    """Same tolerance reached, iteration counts within +-2 of the reference (golden vectors, OMP threads = 1)."""
    A = _mat(orc, name)
    b = np.ones(A.rows)
    dA = dev_csr(pkg, be, A)
    mat = dA if fmt == "csr" else dA.to_sell(32)
    db = be.array(b)

    def run(solver, **kw):
        dx = be.array(np.full(A.rows, 123.0))            # result must not depend on the incoming x
        tag = pkg.SolverTag(**kw).solve(solver, mat, db, dx)
        return tag, dx.download()

    def check(tag, x, key, slack=2):
        it = int(golden["solve/%s/%s/iters" % (name, key)][0])
        assert abs(tag.iters - it) <= slack, (key, tag.iters, it)
        xr = golden["solve/%s/%s/x" % (name, key)]
        assert np.linalg.norm(x - xr) <= 1e-5 * np.linalg.norm(xr)
        assert tag.error < 1e-8
        assert _true_res(A, b, x) < 2e-7

    if name.startswith("lap"):
        tag, x = run("cg", tol=1e-8, max_iterations=1000)
        check(tag, x, "cg_none")
    tag, x = run("bicgstab", tol=1e-8, max_iterations=1000)
    check(tag, x, "bicgstab_none", slack=4)              # BiCGStab: the reference itself moves by a few iterations between runs
    if fmt == "csr":
        tag, x = run("bicgstab", tol=1e-8, max_iterations=1000, precond=1)
        check(tag, x, "bicgstab_jacobi", slack=4)
    tag, x = run("gmres", tol=1e-8, max_iterations=1000, krylov_dim=30)
    check(tag, x, "gmres_pipelined_fixed")
    it_h = int(golden["solve/%s/gmres_identity/iters" % name][0])
    assert tag.iters == -(-it_h // 30) * 30


def test_cg_residual_history_and_monitor(pkg, be, orc):
    """Monitor contract (cg.hpp:174): called once per iteration with the estimate; returning true stops the solver."""
    A = orc.stencil2d(63, 65)
    b = np.ones(A.rows)
    ref = orc.cg(A, b, tol=1e-8, maxit=1000, hist_cap=2000)
    dA = dev_csr(pkg, be, A)
    db, dx = be.array(b), be.zeros(A.rows)
    hist = []
    tag = pkg.SolverTag(tol=1e-8, max_iterations=1000, monitor=lambda xp, est: hist.append(est) or False).solve("cg", dA, db, dx)
    assert len(hist) == tag.iters
    assert abs(tag.iters - ref["iters"]) <= 2
    k = min(len(hist), len(ref["history"]), 100)
    assert np.allclose(hist[:k], ref["history"][:k], rtol=1e-6)
    hist2 = []
    tag = pkg.SolverTag(tol=1e-8, max_iterations=1000, monitor=lambda xp, est: hist2.append(est) or len(hist2) >= 10).solve("cg", dA, db, dx)
    assert len(hist2) == 10 and tag.iters == 10


def test_solver_edge_cases(pkg, be, orc):
    A = orc.stencil2d(31, 33)
    dA = dev_csr(pkg, be, A)
    # zero right-hand side: x = 0, no iterations (cg.hpp:149-151, bicgstab.hpp:140-141)
    db, dx = be.zeros(A.rows), be.array(np.ones(A.rows))
    for s in ("cg", "bicgstab", "gmres"):
        dx.upload(np.ones(A.rows))
        tag = pkg.SolverTag(tol=1e-8, max_iterations=50).solve(s, dA, db, dx)
        assert tag.iters == 0 and np.all(dx.download() == 0.0), s
    # iteration budget exhausted: iters == max_iterations, error reported
    b = np.ones(A.rows)
    db.upload(b)
    ref = orc.cg(A, b, tol=1e-14, maxit=7)
    tag = pkg.SolverTag(tol=1e-14, max_iterations=7).solve("cg", dA, db, dx)
    assert tag.iters == 7 and abs(tag.error - ref["error"]) <= 1e-6 * ref["error"]
    assert np.allclose(dx.download(), ref["x"], rtol=1e-9, atol=1e-12)
    ref = orc.bicgstab(A, b, tol=1e-14, maxit=5)
    tag = pkg.SolverTag(tol=1e-14, max_iterations=5).solve("bicgstab", dA, db, dx)
    assert tag.iters == 5 and np.allclose(dx.download(), ref["x"], rtol=1e-8, atol=1e-12)
    ref = orc.gmres(A, b, tol=1e-14, maxit=25, krylov=10)
    tag = pkg.SolverTag(tol=1e-14, max_iterations=25, krylov_dim=10).solve("gmres", dA, db, dx)
    assert tag.iters == ref["iters"] and np.allclose(dx.download(), ref["x"], rtol=1e-8, atol=1e-12)
    # unsupported combinations fail loudly instead of silently taking another path
    with pytest.raises(pkg.VclError):
        pkg.SolverTag(precond=1, krylov_dim=10).solve("gmres", dA.to_sell(32), db, dx)   # Jacobi needs the CSR matrix (row_info)
    with pytest.raises(pkg.VclError):
        pkg.SolverTag(precond=7).solve("gmres", dA, db, dx)                    # unknown preconditioner id
    with pytest.raises(pkg.VclError):
        pkg.SolverTag(precond=1).solve("cg", dA.to_sell(32), db, dx)           # Jacobi needs the CSR matrix (row_info)


def test_config1_cg_parity_1024(pkg, be, orc):
    """BASELINE config 1: CG on the 2-D 5-point Laplacian 1024x1024, b = 1, tol 1e-8 (reference: 1898 iterations)."""
    orc.set_threads(orc.max_threads())
    A = orc.stencil2d(1024, 1024)
    b = np.ones(A.rows)
    ref = orc.cg(A, b, tol=1e-8, maxit=5000)
    orc.set_threads(1)
    dA = pkg.CsrMatrix.stencil(be, 1024, 1024, 1)
    db, dx = be.array(b), be.zeros(A.rows)
    tag = pkg.SolverTag(tol=1e-8, max_iterations=5000).solve("cg", dA, db, dx)
    assert abs(tag.iters - ref["iters"]) <= 2, (tag.iters, ref["iters"])
    assert abs(tag.iters - 1898) <= 2
    x = dx.download()
    assert _true_res(A, b, x) <= _true_res(A, b, ref["x"]) * (1 + 1e-6) + 1e-8
    tag2 = pkg.SolverTag(tol=1e-8, max_iterations=5000).solve("cg", dA.to_sell(32), db, dx)
    assert abs(tag2.iters - ref["iters"]) <= 2


# ----------------------------------------------------------------------------------------------- GMRES + Jacobi (fused)
def _gmres_jacobi_cases():
    import json, os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gmres_jacobi.json")))


@pytest.mark.parametrize("case", _gmres_jacobi_cases(), ids=lambda c: "%s-m%d-%d" % (c["name"], c["krylov"], c["maxit"]))
def test_gmres_jacobi_fused_vs_reference(pkg, be, orc, case):
    """solve(A, b, gmres_tag, jacobi_precond) on the fused path (pipelined cycle on D^-1 A, per-iteration stopping rule) against
    the reference's Householder path (gmres.hpp:449-631): same iteration count (+-2), estimate and solution."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_gmres_jacobi import vardiag
    nx, ny, nz = case["grid"]; c = case["c"]
    A = vardiag(orc.stencil3d(nx, ny, nz, *c) if nz > 1 else orc.stencil2d(nx, ny, c[0], c[1]))
    b = np.ones(A.rows)
    dA = dev_csr(pkg, be, A)
    db, dx = be.array(b), be.array(np.full(A.rows, 2.0))
    tag = pkg.SolverTag(tol=case["tol"], max_iterations=case["maxit"], krylov_dim=case["krylov"], precond=1).solve("gmres", dA, db, dx)
    assert abs(tag.iters - case["iters"]) <= 2, (tag.iters, case["iters"])
    x = dx.download()
    assert abs(np.linalg.norm(x) - case["x_norm"]) <= 1e-5 * case["x_norm"]
    if case["error"] < case["tol"]:
        assert tag.error < case["tol"]
        assert _true_res(A, b, x) <= case["true_residual"] * 1.5 + 1e-9
    else:
        assert abs(tag.error - case["error"]) <= 0.05 * case["error"]


@pytest.mark.parametrize("precond", [2, 3, 4])
def test_row_scaling_fused_paths(pkg, be, orc, precond):
    """row_scaling (row_scaling.hpp:150-190; inf-/1-/2-norm of the rows) on the fused diagonal-preconditioner paths: the
    scaling vector equals row_info's norms, every solver reaches the direct solution, and on a matrix whose diagonal varies
    over two decades the preconditioned CG needs far fewer iterations (the facade test pins the counts to the reference)."""
    import sys, os
    import scipy.sparse.linalg as sl
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_gmres_jacobi import vardiag
    A = vardiag(orc.stencil2d(40, 36))
    M = A.to_scipy()
    # symmetric variant for CG: D^1/2 L D^1/2 keeps SPD
    s = np.sqrt(10.0 ** np.random.default_rng(11).uniform(-1, 1, A.rows))
    L = orc.stencil2d(40, 36)
    import scipy.sparse as sp
    S = (sp.diags(s) @ L.to_scipy() @ sp.diags(s)).tocsr(); S.sort_indices()
    Asym = ol.CSR(A.rows, A.cols, S.indptr.astype(np.uint32), S.indices.astype(np.uint32), S.data)
    b = np.ones(A.rows)
    for solver, mat, Msp in (("cg", Asym, S), ("bicgstab", A, M), ("gmres", A, M)):
        dA = dev_csr(pkg, be, mat)
        db, dx, dx0 = be.array(b), be.zeros(A.rows), be.zeros(A.rows)
        tag = pkg.SolverTag(tol=1e-9, max_iterations=3000, krylov_dim=30, precond=precond).solve(solver, dA, db, dx)
        exact = sl.spsolve(Msp.tocsc(), b)
        assert np.linalg.norm(dx.download() - exact) <= 1e-6 * np.linalg.norm(exact), (solver, precond)
        if solver == "cg":
            plain = pkg.SolverTag(tol=1e-9, max_iterations=3000).solve(solver, dA, db, dx0)
            assert tag.iters < plain.iters, (tag.iters, plain.iters)
    norms = {2: abs(M).max(axis=1).toarray().ravel(), 3: np.asarray(abs(M).sum(axis=1)).ravel(), 4: np.sqrt(np.asarray(M.multiply(M).sum(axis=1)).ravel())}
    dA = dev_csr(pkg, be, A)
    assert np.allclose(dA.row_info(precond - 2).download(), norms[precond], rtol=1e-14)


def test_cg_persistent_kernel_on_ragged_spd_system(pkg, be, orc):
    """The persistent cooperative CG kernel (small systems) on a matrix that is nothing like a stencil: odd size, empty-ish and
    very long rows (whole-CTA row path), fewer row blocks than CTAs.  Same iteration count (+-2) and solution as the oracle's
    pipelined CG; a monitor (one iteration per launch) must see the same sequence of estimates."""
    import scipy.sparse as sp
    rng = np.random.default_rng(21)
    n = 6001
    lens = rng.integers(0, 12, n); lens[17] = 3000; lens[4000] = 2500
    rows = np.repeat(np.arange(n), lens)
    cols = np.concatenate([rng.choice(n, l, replace=False) for l in lens])
    P = sp.coo_matrix((rng.uniform(-1, 0, rows.size), (rows, cols)), shape=(n, n)).tocsr()
    S = P + P.T
    S = (S + sp.diags(np.asarray(abs(S).sum(axis=1)).ravel() + 1.0)).tocsr()
    S.sum_duplicates(); S.sort_indices()
    A = ol.CSR(n, n, S.indptr.astype(np.uint32), S.indices.astype(np.uint32), S.data)
    b = orc.uniform(n, 3, -1.0, 1.0)
    ref = orc.cg(A, b, tol=1e-10, maxit=500, hist_cap=500)
    dA = dev_csr(pkg, be, A)
    db, dx = be.array(b), be.zeros(n)
    tag = pkg.SolverTag(tol=1e-10, max_iterations=500).solve("cg", dA, db, dx)
    assert abs(tag.iters - ref["iters"]) <= 2, (tag.iters, ref["iters"])
    assert np.linalg.norm(dx.download() - ref["x"]) <= 1e-8 * np.linalg.norm(ref["x"])
    assert _true_res(A, b, dx.download()) < 1e-9
    hist = []
    tag2 = pkg.SolverTag(tol=1e-10, max_iterations=500, monitor=lambda xp, est: hist.append(est) or False).solve("cg", dA, db, dx)
    assert tag2.iters == tag.iters and len(hist) == tag.iters
    m = min(len(hist), len(ref["history"]))
    rel = np.abs(np.array(hist[:m]) - ref["history"][:m]) / ref["history"][:m]
    # rows of 3000 entries are tree-summed on the device and summed sequentially by the oracle: the estimates agree to rounding
    # at first; on this badly scaled system the pipelined recurrence amplifies 1e-16 perturbations by ~1e5 per iteration
    # (measured: 6e-16, 2e-15, 9e-15, 7e-15, 8e-15, 1e-12, 1e-7, 2e-2), the iteration count does not move
    assert rel[:5].max() < 1e-12, rel
    few = pkg.SolverTag(tol=1e-30, max_iterations=37).solve("cg", dA, db, dx)          # budget not a multiple of the batch
    assert few.iters == 37


def test_foreign_row_block_plans(pkg, be, orc):
    """Row-block plans NOT made by ViennaCLCUDAcsr_row_blocks: the reference's own handle3() blocks
    (compressed_matrix.hpp:1152-1188: <= 1024 entries per block, any number of rows), one block holding the whole matrix, and a
    malformed plan.  The library checks a foreign plan once (vcl_plan_ok) and sends the product to the plan-free kernel when the
    plan breaks the limits of the TMA kernel -- the result must be the reference's in every case, and rewriting the plan's
    memory through the C-ABI must drop the cached verdict."""
    rng = np.random.default_rng(5)
    n = 5000
    # rows of 0..3 entries: a reference-style block collects ~680 rows (> 256)
    cnt = rng.integers(0, 4, n)
    rp = np.zeros(n + 1, np.uint32); rp[1:] = np.cumsum(cnt)
    ci = np.concatenate([np.sort(rng.choice(n, c, replace=False)) for c in cnt]).astype(np.uint32)
    va = rng.uniform(-1, 1, int(rp[-1]))
    x = rng.uniform(1, 2, n)
    import scipy.sparse as sp
    y_ref = sp.csr_matrix((va, ci.astype(np.int64), rp.astype(np.int64)), shape=(n, n)) @ x
    dA = pkg.CsrMatrix.from_host(be, n, n, rp, ci, va)
    dx = be.array(x)

    def ref_blocks():                                 # the reference's generate_row_block_information (compressed_matrix.hpp:1152-1188), restated
        blk, acc, i = [0], 0, 0
        while i < n:
            acc += int(rp[i + 1] - rp[i])
            if acc > 1024:
                if i - blk[-1] > 0:
                    blk.append(i); i -= 1             # the current row opens the next batch
                else:
                    blk.append(i + 1)                 # a row longer than the buffer
                acc = 0
            i += 1
        if acc > 0:
            blk.append(n)
        return np.array(blk, np.uint32)

    own = dA.blocks.download()[:dA.nblocks + 1]
    plans = {"own": None, "own_copy": own.copy(), "reference": ref_blocks(), "one_block": np.array([0, n], np.uint32), "pairs": np.arange(0, n + 1, 2, dtype=np.uint32)}
    assert np.diff(plans["reference"].astype(np.int64)).max() > 256
    for name, blk in plans.items():
        dy = be.array(np.full(n, 7.0))
        if blk is not None:
            dA.blocks = be.array(blk); dA.nblocks = len(blk) - 1
        l0 = be.launches()
        dA.spmv(dx, dy)
        dA.spmv(dx, dy)
        assert be.launches() - l0 == (2 if name == "own" else 3), name          # foreign plan: one check launch, once
        assert ol_rel(dy.download(), y_ref) <= 1e-13, name
    # same buffer, new contents written through the C-ABI: the verdict must not survive
    good = plans["pairs"]
    dA.blocks = be.array(good); dA.nblocks = len(good) - 1
    dy = be.zeros(n)
    dA.spmv(dx, dy)
    bad = good.copy(); bad[1:-1] = 0; bad[-2] = n          # same length: [0, 0, ..., 0, n, n] = one huge block in the middle
    dA.blocks.upload(bad)
    dy2 = be.zeros(n)
    dA.spmv(dx, dy2)
    assert ol_rel(dy2.download(), y_ref) <= 1e-13


def ol_rel(a, b):
    import oracle_lib as ol
    return float(ol.rel_err(a, b).max())
