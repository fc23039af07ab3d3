"""Golden vectors for the ELL / HYB / COO formats, produced by the UNMODIFIED reference (oracle/_ref, 1 thread) on the
matrices of reference_vectors.npz:  python tests/golden/make_golden_formats.py  ->  tests/golden/formats_vectors.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402


def vardiag_matrix(o, scale):
    """SPD test system: 2-D 5-point Laplacian 63 x 65 plus diag(scale * ((7919 r) mod 1000) / 999)."""
    import scipy.sparse as sp
    L = o.stencil2d(63, 65)
    S = L.to_scipy().tocsr()
    n = S.shape[0]
    S = (S + sp.diags(scale * (((np.arange(n) * 7919) % 1000) / 999.0))).tocsr()
    S.sort_indices()
    return ol.CSR(n, n, S.indptr.astype(np.uint32), S.indices.astype(np.uint32), S.data)


def main():
    o = ol.oracle(); r = ol.ref()
    r.set_threads(1)
    g = np.load(os.path.join(HERE, "reference_vectors.npz"))
    out = {}
    for name in ["lap2d_13x11", "cd3d_9x8x7", "ragged_200x180", "ragged_97x97"]:
        rows, cols = g[name + "/shape"]
        A = ol.CSR(rows, cols, g[name + "/rp"], g[name + "/ci"], g[name + "/v"])
        x, y0 = g[name + "/x"], g[name + "/y0"]
        E = r.ell_build(A)
        out[name + "/ell/width"] = np.array([E["width"], E["internal_rows"]])
        out[name + "/ell/coords"] = E["coords"]; out[name + "/ell/elements"] = E["elements"]
        out[name + "/ell/y"] = r.ell_spmv(A, x.copy())
        out[name + "/ell/y_ab"] = r.ell_spmv(A, x.copy(), y0.copy(), 1.5, -0.25)
        H = r.hyb_build(A)
        out[name + "/hyb/width"] = np.array([H["width"], H["internal_rows"], H["csr_nnz"]])
        for k in ("ell_coords", "ell_elements", "csr_rows", "csr_cols", "csr_elements"):
            out[name + "/hyb/" + k] = H[k]
        out[name + "/hyb/y"] = r.hyb_spmv(A, x.copy())
        out[name + "/hyb/y_ab"] = r.hyb_spmv(A, x.copy(), y0.copy(), 1.5, -0.25)
        M = r.coo_build(A)
        out[name + "/coo/coords"] = M["coords"]; out[name + "/coo/elements"] = M["elements"]
        out[name + "/coo/y"] = r.coo_spmv(A, x.copy())
        out[name + "/coo/y_ab"] = r.coo_spmv(A, x.copy(), y0.copy(), 1.5, -0.25)
    # solvers on the other formats (pipelined paths, cg.hpp:204-254 overloads): iteration counts at 1 thread
    L = o.stencil2d(63, 65)
    Cd = o.stencil2d(48, 50, 0.5, 0.0)
    for name, A, solvers in (("lap2d_63x65", L, ("cg", "bicgstab")), ("cd2d_48x50", Cd, ("bicgstab",))):
        b = np.ones(A.rows)
        for fmt, fname in ((2, "ell"), (3, "hyb")):
            for solver in solvers:
                res = r.solve(solver, A, b, precond="none", fmt=fmt, tol=1e-8, maxit=1000)
                key = "solve/%s/%s_%s" % (name, solver, fname)
                out[key + "/iters"] = np.array([res["iters"]]); out[key + "/error"] = np.array([res["error"]]); out[key + "/x"] = res["x"]
    # CG with the Jacobi preconditioner: the reference's generic PCG (cg.hpp:257-322) on Laplacian + varying diagonal
    for scale in (0.5, 5.0):
        A = vardiag_matrix(o, scale)
        res = r.solve("cg", A, np.ones(A.rows), precond="jacobi", tol=1e-9, maxit=2000)
        key = "solve/vd2d_63x65_s%g/cg_jacobi" % scale
        out[key + "/iters"] = np.array([res["iters"]]); out[key + "/error"] = np.array([res["error"]]); out[key + "/x"] = res["x"]
    np.savez_compressed(os.path.join(HERE, "formats_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
