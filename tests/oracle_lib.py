"""ctypes access to the CHECKERS under oracle/ (test infrastructure only).

`oracle`  -> oracle/libvcl_oracle.so        plain-C restatement (always buildable)
`ref`     -> oracle/_ref/libvcl_ref.so      the unmodified reference's OpenMP host backend behind ref_shim.cpp
`ref_fix` -> oracle/_ref/libvcl_ref_gmresfix.so   same with the documented 1-line GMRES fix

Nothing in the product package imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
c_int, c_dbl, c_ll = C.c_int, C.c_double, C.c_longlong


def build(force=False):
    """Compile the checkers (gcc only; `ref` only when /root/reference is present)."""
    so = os.path.join(ORACLE_DIR, "libvcl_oracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(ORACLE_DIR, "vcl_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
    ref = os.path.join(ORACLE_DIR, "_ref", "libvcl_ref.so")
    if os.path.isdir("/root/reference/viennacl") and (force or not os.path.exists(ref)
                                                       or os.path.getmtime(ref) < os.path.getmtime(os.path.join(ORACLE_DIR, "ref_shim.cpp"))):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])


class CSR:
    """Host CSR triple with the reference's array layout (u32 row_ptr[rows+1], u32 col[nnz], f64 val[nnz])."""

    def __init__(self, rows, cols, rp, ci, v, dtype=np.float64):
        self.rows, self.cols = int(rows), int(cols)
        self.rp = np.ascontiguousarray(rp, dtype=np.uint32)
        self.ci = np.ascontiguousarray(ci, dtype=np.uint32)
        self.v = np.ascontiguousarray(v, dtype=dtype)
        self.nnz = int(self.rp[-1]) if self.rp.size else 0

    def astype(self, dtype):
        return CSR(self.rows, self.cols, self.rp, self.ci, self.v.astype(dtype), dtype)

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.v, self.ci.astype(np.int64), self.rp.astype(np.int64)), shape=(self.rows, self.cols))


class _Oracle:
    def __init__(self, dtype=np.float64):
        build()
        self.dt = np.dtype(dtype).type
        single = self.dt is np.float32
        fp = np.ctypeslib.ndpointer(dtype=self.dt, flags="C_CONTIGUOUS")
        cr = C.c_float if single else C.c_double        # vreal scalars; tolerances / error / history stay double
        self.lib = lib = C.CDLL(os.path.join(ORACLE_DIR, "libvcl_oracle_f32.so" if single else "libvcl_oracle.so"))
        lib.vclo_gen_stencil2d.restype = c_ll
        lib.vclo_gen_stencil2d.argtypes = [c_int, c_int, cr, cr, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.vclo_gen_stencil3d.restype = c_ll
        lib.vclo_gen_stencil3d.argtypes = [c_int, c_int, c_int, cr, cr, cr, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.vclo_fill_uniform.argtypes = [fp, c_ll, C.c_ulonglong, cr, cr]
        lib.vclo_csr_spmv.argtypes = [c_int, u32p, u32p, fp, fp, c_int, c_int, cr, fp, c_int, c_int, cr]
        lib.vclo_sell_padded_nnz.restype = c_ll
        lib.vclo_sell_padded_nnz.argtypes = [c_int, u32p, c_int]
        lib.vclo_sell_build.argtypes = [c_int, u32p, u32p, fp, c_int, u32p, u32p, u32p, fp]
        lib.vclo_sell_spmv.argtypes = [c_int, c_int, u32p, u32p, u32p, fp, fp, c_int, c_int, cr, fp, c_int, c_int, cr]
        lib.vclo_csr_diag.argtypes = [c_int, u32p, u32p, fp, fp]
        lib.vclo_ell_width.restype = c_int
        lib.vclo_ell_width.argtypes = [c_int, u32p]
        lib.vclo_ell_build.argtypes = [c_int, u32p, u32p, fp, c_int, u32p, fp]
        lib.vclo_ell_spmv.argtypes = [c_int, c_int, u32p, fp, fp, c_int, c_int, cr, fp, c_int, c_int, cr]
        lib.vclo_hyb_width.restype = c_int
        lib.vclo_hyb_width.argtypes = [c_int, c_int, u32p, c_dbl]
        lib.vclo_hyb_tail_nnz.restype = c_ll
        lib.vclo_hyb_tail_nnz.argtypes = [c_int, u32p, c_int]
        lib.vclo_hyb_build.argtypes = [c_int, u32p, u32p, fp, c_int, u32p, fp, u32p, u32p, fp]
        lib.vclo_hyb_spmv.argtypes = [c_int, c_int, u32p, fp, u32p, u32p, fp, fp, c_int, c_int, cr, fp, c_int, c_int, cr]
        lib.vclo_coo_spmv.argtypes = [c_int, c_ll, u32p, fp, fp, cr, fp, cr]
        lib.vclo_norm2.restype = cr
        lib.vclo_norm2.argtypes = [fp, c_ll]
        lib.vclo_inner_prod.restype = cr
        lib.vclo_inner_prod.argtypes = [fp, fp, c_ll]
        ip, dp = C.POINTER(c_int), C.POINTER(c_dbl)
        lib.vclo_cg.argtypes = [c_int, u32p, u32p, fp, fp, fp, c_dbl, c_dbl, c_int, ip, dp, C.c_void_p, c_int, ip]
        lib.vclo_bicgstab.argtypes = lib.vclo_cg.argtypes
        lib.vclo_bicgstab_precond.argtypes = [c_int, u32p, u32p, fp, c_int, fp, fp, c_dbl, c_dbl, c_int, c_int, ip, dp, C.c_void_p, c_int, ip]
        lib.vclo_gmres.argtypes = [c_int, u32p, u32p, fp, fp, fp, c_dbl, c_dbl, c_int, c_int, ip, dp, C.c_void_p, c_int, ip]
        lib.vclo_max_threads.restype = c_int
        lib.vclo_set_threads.argtypes = [c_int]
        if not single:
            lib.vclo_mixed_cg.argtypes = [c_int, u32p, u32p, fp, fp, fp, c_dbl, c_int, C.c_float, ip, dp, ip]

    # -- generators --------------------------------------------------------------------------
    def stencil2d(self, nx, ny, cx=0.0, cy=0.0):
        nnz = self.lib.vclo_gen_stencil2d(nx, ny, cx, cy, None, None, None)
        n = nx * ny
        rp = np.empty(n + 1, np.uint32); ci = np.empty(nnz, np.uint32); v = np.empty(nnz, self.dt)
        self.lib.vclo_gen_stencil2d(nx, ny, cx, cy, rp.ctypes.data, ci.ctypes.data, v.ctypes.data)
        return CSR(n, n, rp, ci, v, self.dt)

    def stencil3d(self, nx, ny, nz, cx=0.0, cy=0.0, cz=0.0):
        nnz = self.lib.vclo_gen_stencil3d(nx, ny, nz, cx, cy, cz, None, None, None)
        n = nx * ny * nz
        rp = np.empty(n + 1, np.uint32); ci = np.empty(nnz, np.uint32); v = np.empty(nnz, self.dt)
        self.lib.vclo_gen_stencil3d(nx, ny, nz, cx, cy, cz, rp.ctypes.data, ci.ctypes.data, v.ctypes.data)
        return CSR(n, n, rp, ci, v, self.dt)

    def uniform(self, n, seed=42, lo=0.0, hi=1.0):
        x = np.empty(n, self.dt)
        self.lib.vclo_fill_uniform(x, n, seed, lo, hi)
        return x

    def set_threads(self, n):
        self.lib.vclo_set_threads(n)

    def max_threads(self):
        return self.lib.vclo_max_threads()

    # -- SpMV ----------------------------------------------------------------------------------
    def csr_spmv(self, A, x, y=None, alpha=1.0, beta=0.0, offx=0, incx=1, offy=0, incy=1):
        if y is None:
            y = np.zeros(offy + A.rows * incy, self.dt)
        self.lib.vclo_csr_spmv(A.rows, A.rp, A.ci, A.v, x, offx, incx, alpha, y, offy, incy, beta)
        return y

    def sell_build(self, A, Cs=32):
        nb = (A.rows - 1) // Cs + 1 if A.rows > 0 else 0
        tot = self.lib.vclo_sell_padded_nnz(A.rows, A.rp, Cs)
        cpb = np.zeros(max(nb, 1), np.uint32); bs = np.zeros(max(nb, 1), np.uint32)
        ci = np.zeros(max(tot, 1), np.uint32); el = np.zeros(max(tot, 1), self.dt)
        self.lib.vclo_sell_build(A.rows, A.rp, A.ci, A.v, Cs, cpb, bs, ci, el)
        return dict(rows=A.rows, cols=A.cols, C=Cs, nb=nb, padded_nnz=int(tot), cols_per_block=cpb[:nb], block_start=bs[:nb],
                    col_idx=ci[:tot], elements=el[:tot])

    def sell_spmv(self, S, x, y=None, alpha=1.0, beta=0.0, offx=0, incx=1, offy=0, incy=1):
        if y is None:
            y = np.zeros(offy + S["rows"] * incy, self.dt)
        pad = lambda a, dt: np.ascontiguousarray(a if a.size else np.zeros(1, dt))
        self.lib.vclo_sell_spmv(S["rows"], S["C"], pad(S["cols_per_block"], np.uint32), pad(S["block_start"], np.uint32),
                                pad(S["col_idx"], np.uint32), pad(S["elements"], self.dt), x, offx, incx, alpha, y, offy, incy, beta)
        return y

    def csr_diag(self, A):
        d = np.empty(A.rows, self.dt)
        self.lib.vclo_csr_diag(A.rows, A.rp, A.ci, A.v, d)
        return d

    # -- COO (coordinate_matrix.hpp:47-102) ----------------------------------------------------
    def coo_build(self, A):
        rows = np.repeat(np.arange(A.rows, dtype=np.uint32), np.diff(A.rp.astype(np.int64)))
        coords = np.empty(2 * A.nnz, np.uint32)
        coords[0::2] = rows; coords[1::2] = A.ci
        return dict(rows=A.rows, cols=A.cols, nnz=A.nnz, coords=coords, elements=A.v.copy())

    def coo_spmv(self, M, x, y=None, alpha=1.0, beta=0.0):
        if y is None:
            y = np.zeros(M["rows"], self.dt)
        pad = lambda a, dt: np.ascontiguousarray(a if a.size else np.zeros(2, dt))
        self.lib.vclo_coo_spmv(M["rows"], M["nnz"], pad(M["coords"], np.uint32), pad(M["elements"], self.dt), x, alpha, y, beta)
        return y

    # -- ELL / HYB (ell_matrix.hpp:122-166, hyb_matrix.hpp:127-214) ---------------------------
    def ell_build(self, A):
        w = self.lib.vclo_ell_width(A.rows, A.rp)
        tot = max(A.rows * w, 1)
        co = np.zeros(tot, np.uint32); el = np.zeros(tot, self.dt)
        self.lib.vclo_ell_build(A.rows, A.rp, A.ci, A.v, w, co, el)
        return dict(rows=A.rows, cols=A.cols, width=w, internal_rows=A.rows, coords=co[:A.rows * w], elements=el[:A.rows * w])

    def ell_spmv(self, E, x, y=None, alpha=1.0, beta=0.0, offx=0, incx=1, offy=0, incy=1):
        if y is None:
            y = np.zeros(offy + E["rows"] * incy, self.dt)
        pad = lambda a, dt: np.ascontiguousarray(a if a.size else np.zeros(1, dt))
        self.lib.vclo_ell_spmv(E["rows"], E["width"], pad(E["coords"], np.uint32), pad(E["elements"], self.dt),
                               x, offx, incx, alpha, y, offy, incy, beta)
        return y

    def hyb_build(self, A, threshold=0.8):
        w = self.lib.vclo_hyb_width(A.rows, A.cols, A.rp, threshold)
        tn = int(self.lib.vclo_hyb_tail_nnz(A.rows, A.rp, w))
        tot = max(A.rows * w, 1)
        co = np.zeros(tot, np.uint32); el = np.zeros(tot, self.dt)
        cr = np.zeros(A.rows + 1, np.uint32); cc = np.zeros(tn, np.uint32); ce = np.zeros(tn, self.dt)
        self.lib.vclo_hyb_build(A.rows, A.rp, A.ci, A.v, w, co, el, cr, cc, ce)
        return dict(rows=A.rows, cols=A.cols, width=w, internal_rows=A.rows, ell_coords=co[:A.rows * w], ell_elements=el[:A.rows * w],
                    csr_rows=cr, csr_cols=cc, csr_elements=ce, csr_nnz=tn)

    def hyb_spmv(self, H, x, y=None, alpha=1.0, beta=0.0, offx=0, incx=1, offy=0, incy=1):
        if y is None:
            y = np.zeros(offy + H["rows"] * incy, self.dt)
        pad = lambda a, dt: np.ascontiguousarray(a if a.size else np.zeros(1, dt))
        self.lib.vclo_hyb_spmv(H["rows"], H["width"], pad(H["ell_coords"], np.uint32), pad(H["ell_elements"], self.dt),
                               H["csr_rows"], H["csr_cols"], H["csr_elements"], x, offx, incx, alpha, y, offy, incy, beta)
        return y

    def norm2(self, x):
        return self.lib.vclo_norm2(np.ascontiguousarray(x), x.size)

    def inner_prod(self, x, y):
        return self.lib.vclo_inner_prod(np.ascontiguousarray(x), np.ascontiguousarray(y), x.size)

    # -- solvers -------------------------------------------------------------------------------
    def _run(self, fn, A, b, extra, hist_cap):
        x = np.zeros(A.rows, self.dt)
        it, err, hl = c_int(0), c_dbl(0), c_int(0)
        hist = np.zeros(max(hist_cap, 1), np.float64)
        fn(A.rows, A.rp, A.ci, A.v, *extra(b, x), C.byref(it), C.byref(err), hist.ctypes.data, hist_cap, C.byref(hl))
        return dict(x=x, iters=it.value, error=err.value, history=hist[:min(hl.value, hist_cap)].copy())

    def cg(self, A, b, tol=1e-8, maxit=300, abs_tol=0.0, hist_cap=0):
        return self._run(self.lib.vclo_cg, A, b, lambda b, x: (b, x, tol, abs_tol, maxit), hist_cap)

    def bicgstab(self, A, b, tol=1e-8, maxit=400, abs_tol=0.0, hist_cap=0):
        return self._run(self.lib.vclo_bicgstab, A, b, lambda b, x: (b, x, tol, abs_tol, maxit), hist_cap)

    def bicgstab_precond(self, A, b, precond=1, tol=1e-8, maxit=400, abs_tol=0.0, restart_every=200, hist_cap=0):
        x = np.zeros(A.rows, self.dt)
        it, err, hl = c_int(0), c_dbl(0), c_int(0)
        hist = np.zeros(max(hist_cap, 1), np.float64)
        self.lib.vclo_bicgstab_precond(A.rows, A.rp, A.ci, A.v, precond, b, x, tol, abs_tol, maxit, restart_every,
                                       C.byref(it), C.byref(err), hist.ctypes.data, hist_cap, C.byref(hl))
        return dict(x=x, iters=it.value, error=err.value, history=hist[:min(hl.value, hist_cap)].copy())

    def mixed_cg(self, A, b, tol=1e-8, maxit=300, inner_tol=1e-2):
        """mixed_precision_cg.hpp:95-186 (double system, float inner iterations); double oracle only."""
        x = np.zeros(A.rows, np.float64)
        it, err, up = c_int(0), c_dbl(0), c_int(0)
        self.lib.vclo_mixed_cg(A.rows, A.rp, A.ci, A.v, b, x, tol, maxit, inner_tol, C.byref(it), C.byref(err), C.byref(up))
        return dict(x=x, iters=it.value, error=err.value, outer_updates=up.value)

    def gmres(self, A, b, tol=1e-10, maxit=300, krylov=20, abs_tol=0.0, hist_cap=0):
        return self._run(self.lib.vclo_gmres, A, b, lambda b, x: (b, x, tol, abs_tol, maxit, krylov), hist_cap)


class _Ref:
    """The reference itself (OpenMP host backend).  Present only where oracle/_ref/*.so was built/prebuilt."""

    def __init__(self, fixed=False, dtype=np.float64):
        build()
        self.dt = np.dtype(dtype).type
        single = self.dt is np.float32
        fp = np.ctypeslib.ndpointer(dtype=self.dt, flags="C_CONTIGUOUS")
        cr = C.c_float if single else C.c_double        # real_t scalars; tolerances / error / history / seconds stay double
        self.ct = cr
        name = ("libvcl_ref_gmresfix" if fixed else "libvcl_ref") + ("_f32.so" if single else ".so")
        path = os.path.join(ORACLE_DIR, "_ref", name)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = lib = C.CDLL(path)
        ip, dp = C.POINTER(c_int), C.POINTER(c_dbl)
        lib.vclref_max_threads.restype = c_int
        lib.vclref_set_threads.argtypes = [c_int]
        lib.vclref_csr_spmv.argtypes = [c_int, c_int, c_int, u32p, u32p, fp, fp, c_int, c_int, c_int, cr,
                                        fp, c_int, c_int, c_int, cr, c_int]
        lib.vclref_sell_build.argtypes = [c_int, c_int, u32p, u32p, fp, c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), ip, C.POINTER(c_ll)]
        lib.vclref_free.argtypes = [C.c_void_p]
        lib.vclref_sell_spmv.argtypes = [c_int, c_int, u32p, u32p, fp, c_int, fp, cr, fp, cr]
        lib.vclref_csr_diag.argtypes = [c_int, c_int, c_int, u32p, u32p, fp, fp]
        vpp = C.POINTER(C.c_void_p)
        if hasattr(lib, "vclref_ell_build"):
            lib.vclref_ell_build.argtypes = [c_int, c_int, u32p, u32p, fp, vpp, vpp, ip, ip]
            lib.vclref_ell_spmv.argtypes = [c_int, c_int, u32p, u32p, fp, fp, cr, fp, cr]
            lib.vclref_hyb_build.argtypes = [c_int, c_int, u32p, u32p, fp, vpp, vpp, ip, ip, vpp, vpp, vpp, ip]
            lib.vclref_hyb_spmv.argtypes = [c_int, c_int, u32p, u32p, fp, fp, cr, fp, cr]
        if hasattr(lib, "vclref_coo_build"):
            lib.vclref_coo_build.argtypes = [c_int, c_int, u32p, u32p, fp, vpp, vpp, ip]
            lib.vclref_coo_spmv.argtypes = [c_int, c_int, u32p, u32p, fp, fp, cr, fp, cr]
        lib.vclref_norm2.restype = cr
        lib.vclref_norm2.argtypes = [fp, c_int]
        lib.vclref_inner_prod.restype = cr
        lib.vclref_inner_prod.argtypes = [fp, fp, c_int]
        lib.vclref_solve.argtypes = [c_int, c_int, c_int, c_int, c_int, u32p, u32p, fp, fp, fp,
                                     c_dbl, c_dbl, c_int, c_int, c_int, ip, dp, C.c_void_p, c_int, ip, dp]
        if not single and hasattr(lib, "vclref_mixed_cg"):
            lib.vclref_mixed_cg.argtypes = [c_int, c_int, u32p, u32p, fp, fp, fp, c_dbl, c_int, C.c_float, ip, dp]
        lib.vclref_time_csr_spmv.restype = c_dbl
        lib.vclref_time_csr_spmv.argtypes = [c_int, c_int, c_int, u32p, u32p, fp, fp, fp, c_int]

    def set_threads(self, n):
        self.lib.vclref_set_threads(n)

    def max_threads(self):
        return self.lib.vclref_max_threads()

    def csr_spmv(self, A, x, y=None, alpha=1.0, beta=0.0, offx=0, incx=1, offy=0, incy=1, mode=0, nx=None, ny=None):
        if nx is None:
            nx = A.cols
        if ny is None:
            ny = A.rows
        if y is None:
            y = np.zeros(offy + A.rows * incy, self.dt)
        rc = self.lib.vclref_csr_spmv(A.rows, A.cols, A.nnz, A.rp, A.ci, A.v, x, offx, incx, nx, alpha, y, offy, incy, ny, beta, mode)
        assert rc == 0
        return y

    def sell_build(self, A, Cs=32):
        p = [C.c_void_p() for _ in range(4)]
        nb, tot = c_int(0), c_ll(0)
        self.lib.vclref_sell_build(A.rows, A.cols, A.rp, A.ci, A.v, Cs, *[C.byref(q) for q in p], C.byref(nb), C.byref(tot))
        def grab(ptr, n, ct, dt):
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(max(n, 1),))[:n].astype(dt, copy=True)
            return a
        out = dict(rows=A.rows, cols=A.cols, C=Cs, nb=nb.value, padded_nnz=int(tot.value),
                   cols_per_block=grab(p[0], nb.value, C.c_uint32, np.uint32), block_start=grab(p[1], nb.value, C.c_uint32, np.uint32),
                   col_idx=grab(p[2], tot.value, C.c_uint32, np.uint32), elements=grab(p[3], tot.value, self.ct, self.dt))
        for q in p:
            self.lib.vclref_free(q)
        return out

    def sell_spmv(self, A, x, y=None, alpha=1.0, beta=0.0, Cs=32):
        if y is None:
            y = np.zeros(A.rows, self.dt)
        rc = self.lib.vclref_sell_spmv(A.rows, A.cols, A.rp, A.ci, A.v, Cs, x, alpha, y, beta)
        if rc == 3:
            raise ValueError("reference host SELL over-reads when rows % C == 0 (SURVEY 8c-2)")
        return y

    def csr_diag(self, A):
        d = np.empty(A.rows, self.dt)
        self.lib.vclref_csr_diag(A.rows, A.cols, A.nnz, A.rp, A.ci, A.v, d)
        return d

    @staticmethod
    def _grab(ptr, n, ct, dt):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(max(n, 1),))[:n].astype(dt, copy=True)

    def ell_build(self, A):
        p = [C.c_void_p() for _ in range(2)]
        w, ir = c_int(0), c_int(0)
        self.lib.vclref_ell_build(A.rows, A.cols, A.rp, A.ci, A.v, C.byref(p[0]), C.byref(p[1]), C.byref(w), C.byref(ir))
        tot = w.value * ir.value
        out = dict(rows=A.rows, cols=A.cols, width=w.value, internal_rows=ir.value,
                   coords=self._grab(p[0], tot, C.c_uint32, np.uint32), elements=self._grab(p[1], tot, self.ct, self.dt))
        for q in p:
            self.lib.vclref_free(q)
        return out

    def ell_spmv(self, A, x, y=None, alpha=1.0, beta=0.0):
        if y is None:
            y = np.zeros(A.rows, self.dt)
        self.lib.vclref_ell_spmv(A.rows, A.cols, A.rp, A.ci, A.v, x, alpha, y, beta)
        return y

    def coo_build(self, A):
        p = [C.c_void_p() for _ in range(2)]
        n = c_int(0)
        self.lib.vclref_coo_build(A.rows, A.cols, A.rp, A.ci, A.v, C.byref(p[0]), C.byref(p[1]), C.byref(n))
        out = dict(rows=A.rows, cols=A.cols, nnz=n.value, coords=self._grab(p[0], 2 * n.value, C.c_uint32, np.uint32),
                   elements=self._grab(p[1], n.value, self.ct, self.dt))
        for q in p:
            self.lib.vclref_free(q)
        return out

    def coo_spmv(self, A, x, y=None, alpha=1.0, beta=0.0):
        if y is None:
            y = np.zeros(A.rows, self.dt)
        self.lib.vclref_coo_spmv(A.rows, A.cols, A.rp, A.ci, A.v, x, alpha, y, beta)
        return y

    def hyb_build(self, A):
        p = [C.c_void_p() for _ in range(5)]
        w, ir, cn = c_int(0), c_int(0), c_int(0)
        self.lib.vclref_hyb_build(A.rows, A.cols, A.rp, A.ci, A.v, C.byref(p[0]), C.byref(p[1]), C.byref(w), C.byref(ir),
                                  C.byref(p[2]), C.byref(p[3]), C.byref(p[4]), C.byref(cn))
        tot = w.value * ir.value
        out = dict(rows=A.rows, cols=A.cols, width=w.value, internal_rows=ir.value, csr_nnz=cn.value,
                   ell_coords=self._grab(p[0], tot, C.c_uint32, np.uint32), ell_elements=self._grab(p[1], tot, self.ct, self.dt),
                   csr_rows=self._grab(p[2], A.rows + 1, C.c_uint32, np.uint32), csr_cols=self._grab(p[3], cn.value, C.c_uint32, np.uint32),
                   csr_elements=self._grab(p[4], cn.value, self.ct, self.dt))
        for q in p:
            self.lib.vclref_free(q)
        return out

    def hyb_spmv(self, A, x, y=None, alpha=1.0, beta=0.0):
        if y is None:
            y = np.zeros(A.rows, self.dt)
        self.lib.vclref_hyb_spmv(A.rows, A.cols, A.rp, A.ci, A.v, x, alpha, y, beta)
        return y

    def norm2(self, x):
        return self.lib.vclref_norm2(np.ascontiguousarray(x), x.size)

    def inner_prod(self, x, y):
        return self.lib.vclref_inner_prod(np.ascontiguousarray(x), np.ascontiguousarray(y), x.size)

    SOLVERS = dict(cg=0, bicgstab=1, gmres=2)
    PRECONDS = dict(none=0, jacobi=1, identity=2)

    def solve(self, solver, A, b, precond="none", fmt=0, tol=1e-8, abs_tol=0.0, maxit=300, krylov=20, restart_every=200, hist_cap=0):
        x = np.zeros(A.rows, self.dt)
        it, err, hl, sec = c_int(0), c_dbl(0), c_int(0), c_dbl(0)
        hist = np.zeros(max(hist_cap, 1), np.float64)
        rc = self.lib.vclref_solve(self.SOLVERS[solver], self.PRECONDS[precond], fmt, A.rows, A.nnz, A.rp, A.ci, A.v, b, x,
                                   tol, abs_tol, maxit, krylov, restart_every, C.byref(it), C.byref(err),
                                   hist.ctypes.data if hist_cap else None, hist_cap, C.byref(hl) if hist_cap else None, C.byref(sec))
        if rc != 0:
            raise RuntimeError("vclref_solve rc=%d" % rc)
        return dict(x=x, iters=it.value, error=err.value, history=hist[:min(hl.value, hist_cap)].copy(), seconds=sec.value)

    def mixed_cg(self, A, b, tol=1e-8, maxit=300, inner_tol=1e-2):
        x = np.zeros(A.rows, np.float64)
        it, err = c_int(0), c_dbl(0)
        rc = self.lib.vclref_mixed_cg(A.rows, A.nnz, A.rp, A.ci, A.v, b, x, tol, maxit, inner_tol, C.byref(it), C.byref(err))
        assert rc == 0
        return dict(x=x, iters=it.value, error=err.value)

    def time_csr_spmv(self, A, x, reps):
        y = np.zeros(A.rows, self.dt)
        return self.lib.vclref_time_csr_spmv(A.rows, A.cols, A.nnz, A.rp, A.ci, A.v, x, y, reps)


class _RefCuda:
    """The reference's OWN CUDA backend compiled for sm_100 (oracle/ref_cuda_shim.cu): bench.py's `legacy_cuda_baseline`."""

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "_ref", "libvcl_ref_cuda.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = lib = C.CDLL(path)
        ip, dp = C.POINTER(c_int), C.POINTER(c_dbl)
        lib.vclrefcuda_available.restype = c_int
        lib.vclrefcuda_spmv.argtypes = [c_int, c_int, c_int, c_int, u32p, u32p, f64p, f64p, f64p, c_int, dp]
        lib.vclrefcuda_solve.argtypes = [c_int, c_int, c_int, c_int, u32p, u32p, f64p, f64p, f64p, c_dbl, c_int, c_int, ip, dp, dp]

    def available(self):
        return bool(self.lib.vclrefcuda_available())

    def spmv(self, A, x, reps=10, fmt="csr"):
        y = np.zeros(A.rows)
        sec = c_dbl(0)
        rc = self.lib.vclrefcuda_spmv(0 if fmt == "csr" else 1, A.rows, A.cols, A.nnz, A.rp, A.ci, A.v, np.ascontiguousarray(x), y, reps, C.byref(sec))
        if rc != 0:
            raise RuntimeError("vclrefcuda_spmv rc=%d" % rc)
        return y, sec.value

    def solve(self, solver, A, b, precond="none", tol=1e-8, maxit=300, krylov=20):
        x = np.zeros(A.rows)
        it, err, sec = c_int(0), c_dbl(0), c_dbl(0)
        rc = self.lib.vclrefcuda_solve(_Ref.SOLVERS[solver], 1 if precond == "jacobi" else 0, A.rows, A.nnz, A.rp, A.ci, A.v,
                                       np.ascontiguousarray(b), x, tol, maxit, krylov, C.byref(it), C.byref(err), C.byref(sec))
        if rc != 0:
            raise RuntimeError("vclrefcuda_solve rc=%d" % rc)
        return dict(x=x, iters=it.value, error=err.value, seconds=sec.value)


_cache = {}


def ref_cuda():
    """None when the shim is not built or no GPU is visible."""
    if "rc" not in _cache:
        try:
            r = _RefCuda()
            _cache["rc"] = r if r.available() else None
        except (FileNotFoundError, OSError):
            _cache["rc"] = None
    return _cache["rc"]


def oracle(dtype=np.float64):
    key = ("o", np.dtype(dtype).name)
    if key not in _cache:
        _cache[key] = _Oracle(dtype)
    return _cache[key]


def ref(fixed=False, dtype=np.float64):
    key = ("rf" if fixed else "r", np.dtype(dtype).name)
    if key not in _cache:
        _cache[key] = _Ref(fixed, dtype)
    return _cache[key]


def have_ref(dtype=np.float64):
    try:
        ref(dtype=dtype)
        return True
    except (FileNotFoundError, OSError):
        return False


def rel_err(a, b):
    """Per-entry relative difference, the metric of the reference's tests/src/sparse.cpp:66-101."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    m = np.maximum(np.abs(a), np.abs(b))
    d = np.abs(a - b)
    out = np.zeros_like(d)
    nz = m > 0
    out[nz] = d[nz] / m[nz]
    return out
