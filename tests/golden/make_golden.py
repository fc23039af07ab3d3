"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libvcl_ref.so, OpenMP host backend,
OMP threads = 1 so that reductions are deterministic).  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

The fixtures are small and committed; the GPU box never needs /root/reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402


def random_csr(rows, cols, seed, empty_every=7, long_row=None):
    """Ragged CSR: empty rows, a very long row, unsorted-free ascending columns."""
    rng = np.random.default_rng(seed)
    rp = [0]; ci = []; v = []
    for r in range(rows):
        if empty_every and r % empty_every == 3:
            n = 0
        elif long_row is not None and r == long_row:
            n = cols
        else:
            n = int(rng.integers(1, min(cols, 12) + 1))
        c = np.sort(rng.choice(cols, size=n, replace=False))
        ci.extend(c.tolist()); v.extend(rng.uniform(-1.0, 1.0, size=n).tolist())
        rp.append(len(ci))
    return ol.CSR(rows, cols, np.array(rp, np.uint32), np.array(ci, np.uint32), np.array(v, np.float64))


def main():
    o = ol.oracle(); r = ol.ref(); rf = ol.ref(True)
    r.set_threads(1); rf.set_threads(1)
    out = {}

    # ---- SpMV cases (structure of the reference's tests/src/sparse.cpp: plain, +=, -=, aliasing, strided) ----
    mats = {
        "lap2d_13x11": o.stencil2d(13, 11),
        "cd3d_9x8x7": o.stencil3d(9, 8, 7, 0.5, 0.25, 0.125),
        "ragged_200x180": random_csr(200, 180, 7, long_row=50),
        "ragged_97x97": random_csr(97, 97, 11, empty_every=5),
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
        # strided / ranged views: x read at 3 + 2*col, y written at 1 + 3*row (cf. sparse.cpp:163-200)
        xs = o.uniform(3 + 2 * A.cols, 5, 1.0, 2.0)
        ys = o.uniform(1 + 3 * A.rows, 6, -1.0, 1.0)
        out[name + "/xs"] = xs; out[name + "/ys0"] = ys
        out[name + "/ys"] = r.csr_spmv(A, xs.copy(), ys.copy(), offx=3, incx=2, nx=A.cols, offy=1, incy=3, ny=A.rows, mode=0)
        if A.rows == A.cols:
            xa = x.copy()
            r.csr_spmv(A, xa, y0.copy(), mode=4)   # x = A*x
            out[name + "/x_alias"] = xa
        out[name + "/diag"] = r.csr_diag(A)
        S = r.sell_build(A, 32)
        for k in ("cols_per_block", "block_start", "col_idx", "elements"):
            out[name + "/sell32/" + k] = S[k]
        if A.rows % 32 != 0:
            out[name + "/sell32/y"] = r.sell_spmv(A, x.copy(), alpha=1.0, beta=0.0)
            out[name + "/sell32/y_ab"] = r.sell_spmv(A, x.copy(), y0.copy(), alpha=1.5, beta=-0.25)

    # ---- solver cases (the reference has no solver tests; these pin its behaviour at 1 thread) ----
    L = o.stencil2d(63, 65)
    Cd = o.stencil2d(48, 50, 0.5, 0.0)
    C3 = o.stencil3d(11, 10, 9, 0.5, 0.25, 0.125)
    for name, A in (("lap2d_63x65", L), ("cd2d_48x50", Cd), ("cd3d_11x10x9", C3)):
        b = np.ones(A.rows)
        out["solve/" + name + "/gen"] = np.array([0])
        runs = [("bicgstab", "none", dict(tol=1e-8, maxit=1000)),
                ("bicgstab", "jacobi", dict(tol=1e-8, maxit=1000)),
                ("gmres", "identity", dict(tol=1e-8, maxit=1000, krylov=30))]
        if name.startswith("lap"):
            runs.insert(0, ("cg", "none", dict(tol=1e-8, maxit=1000)))
        for solver, pre, kw in runs:
            res = r.solve(solver, A, b, precond=pre, hist_cap=2000, **kw)
            key = "solve/%s/%s_%s" % (name, solver, pre)
            out[key + "/iters"] = np.array([res["iters"]]); out[key + "/error"] = np.array([res["error"]])
            out[key + "/x"] = res["x"]; out[key + "/history"] = res["history"]
        res = rf.solve("gmres", A, b, precond="none", tol=1e-8, maxit=1000, krylov=30, hist_cap=2000)
        key = "solve/%s/gmres_pipelined_fixed" % name
        out[key + "/iters"] = np.array([res["iters"]]); out[key + "/error"] = np.array([res["error"]])
        out[key + "/x"] = res["x"]; out[key + "/history"] = res["history"]

    # ---- BLAS-1 ----
    a = o.uniform(10007, 3, -1.0, 1.0); c = o.uniform(10007, 4, -1.0, 1.0)
    out["blas1/a"] = a; out["blas1/c"] = c
    out["blas1/norm2"] = np.array([r.norm2(a)]); out["blas1/inner"] = np.array([r.inner_prod(a, c)])

    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
