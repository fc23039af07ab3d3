"""Generates tests/golden/reference_vectors_f32.npz from the UNMODIFIED reference instantiated for NumericT = float
(oracle/_ref/libvcl_ref_f32.so: ref_shim.cpp built with -DVCLREF_F32; OpenMP host backend, 1 thread).
Run in the build container (needs /root/reference):

    python tests/golden/make_golden_f32.py

The single-precision fixtures pin oracle/libvcl_oracle_f32.so (tests/test_oracle.py) and the ViennaCLCUDAS... entry points
(tests/test_gpu_float.py).  Tolerances / iteration limits of the solver runs are chosen for float (tol 1e-5).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
from make_golden import random_csr  # noqa: E402

F = np.float32


def main():
    o = ol.oracle(F); r = ol.ref(dtype=F); rf = ol.ref(True, dtype=F)
    r.set_threads(1); rf.set_threads(1)
    out = {}
    rng = np.random.default_rng(2024)

    def perturbed(A):
        # stencil values are small integers: scale them so that products and sums actually round in float
        B = A.astype(F)
        B.v[:] = (B.v * rng.uniform(0.5, 1.5, B.v.size)).astype(F)
        return B

    mats = {
        "lap2d_13x11": perturbed(o.stencil2d(13, 11)),
        "cd3d_9x8x7": perturbed(o.stencil3d(9, 8, 7, 0.5, 0.25, 0.125)),
        "ragged_200x180": random_csr(200, 180, 7, long_row=50).astype(F),
        "ragged_97x97": random_csr(97, 97, 11, empty_every=5).astype(F),
    }
    for name, A in mats.items():
        x = o.uniform(A.cols, 1234, 1.0, 2.0)
        y0 = o.uniform(A.rows, 99, -1.0, 1.0)
        out[name + "/rp"] = A.rp; out[name + "/ci"] = A.ci; out[name + "/v"] = A.v
        out[name + "/shape"] = np.array([A.rows, A.cols])
        out[name + "/x"] = x; out[name + "/y0"] = y0
        out[name + "/y_assign"] = r.csr_spmv(A, x.copy(), y0.copy(), mode=1)
        out[name + "/y_add"] = r.csr_spmv(A, x.copy(), y0.copy(), mode=2)
        out[name + "/y_sub"] = r.csr_spmv(A, x.copy(), y0.copy(), mode=3)
        out[name + "/y_ab"] = r.csr_spmv(A, x.copy(), y0.copy(), alpha=1.5, beta=-0.25, mode=0)
        xs = o.uniform(3 + 2 * A.cols, 5, 1.0, 2.0)
        ys = o.uniform(1 + 3 * A.rows, 6, -1.0, 1.0)
        out[name + "/xs"] = xs; out[name + "/ys0"] = ys
        out[name + "/ys"] = r.csr_spmv(A, xs.copy(), ys.copy(), offx=3, incx=2, nx=A.cols, offy=1, incy=3, ny=A.rows, mode=0)
        out[name + "/diag"] = r.csr_diag(A)
        S = r.sell_build(A, 32)
        for k in ("cols_per_block", "block_start", "col_idx", "elements"):
            out[name + "/sell32/" + k] = S[k]
        if A.rows % 32 != 0:
            out[name + "/sell32/y"] = r.sell_spmv(A, x.copy(), alpha=1.0, beta=0.0)
            out[name + "/sell32/y_ab"] = r.sell_spmv(A, x.copy(), y0.copy(), alpha=1.5, beta=-0.25)
        for fmt, fn in (("ell", r.ell_spmv), ("hyb", r.hyb_spmv), ("coo", r.coo_spmv)):
            out[name + "/" + fmt + "/y"] = fn(A, x.copy())
            out[name + "/" + fmt + "/y_ab"] = fn(A, x.copy(), y0.copy(), alpha=1.5, beta=-0.25)

    # ---- solver runs in float (iteration counts and solutions of the reference at 1 thread) ----
    L = o.stencil2d(63, 65)
    Cd = o.stencil2d(48, 50, 0.5, 0.0)
    for name, A in (("lap2d_63x65", L), ("cd2d_48x50", Cd)):
        b = np.ones(A.rows, F)
        runs = [("bicgstab", "none", dict(tol=1e-5, maxit=1000)),
                ("bicgstab", "jacobi", dict(tol=1e-5, maxit=1000)),
                ("gmres", "identity", dict(tol=1e-5, maxit=1200, krylov=30))]
        if name.startswith("lap"):
            runs.insert(0, ("cg", "none", dict(tol=1e-5, maxit=1000)))
            runs.insert(1, ("cg", "jacobi", dict(tol=1e-5, maxit=1000)))
        for solver, pre, kw in runs:
            res = r.solve(solver, A, b, precond=pre, hist_cap=2000, **kw)
            key = "solve/%s/%s_%s" % (name, solver, pre)
            out[key + "/iters"] = np.array([res["iters"]]); out[key + "/error"] = np.array([res["error"]])
            out[key + "/x"] = res["x"]
            print(key, res["iters"], res["error"])
        res = rf.solve("gmres", A, b, precond="none", tol=1e-5, maxit=1200, krylov=30, hist_cap=2000)
        key = "solve/%s/gmres_pipelined_fixed" % name
        out[key + "/iters"] = np.array([res["iters"]]); out[key + "/error"] = np.array([res["error"]])
        out[key + "/x"] = res["x"]
        print(key, res["iters"], res["error"])

    a = o.uniform(10007, 3, -1.0, 1.0); c = o.uniform(10007, 4, -1.0, 1.0)
    out["blas1/a"] = a; out["blas1/c"] = c
    out["blas1/norm2"] = np.array([r.norm2(a)], F); out["blas1/inner"] = np.array([r.inner_prod(a, c)], F)

    path = os.path.join(HERE, "reference_vectors_f32.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
