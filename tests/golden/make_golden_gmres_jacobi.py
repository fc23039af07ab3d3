"""Generates tests/golden/gmres_jacobi.json: the UNMODIFIED reference's GMRES(m) with jacobi_precond (its Householder path,
gmres.hpp:449-631; host backend, 1 thread) on small systems -- iteration counts, error estimates, solution norms.

    python tests/golden/make_golden_gmres_jacobi.py      (build container only: needs /root/reference)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

CASES = [("cd2d_48x50", (48, 50, 1), (0.5, 0.0, 0.0), 1e-8, 1000, 30), ("cd3d_11x10x9", (11, 10, 9), (0.5, 0.25, 0.125), 1e-8, 1000, 30),
         ("lap2d_63x65", (63, 65, 1), (0.0, 0.0, 0.0), 1e-8, 1000, 30), ("cd3d_24x20x18", (24, 20, 18), (0.5, 0.25, 0.125), 1e-8, 600, 20),
         ("cd2d_48x50", (48, 50, 1), (0.5, 0.0, 0.0), 1e-6, 45, 20)]


def vardiag(A, seed=11):
    """scale the rows so that the diagonal varies over two decades: Jacobi then really changes the Krylov space"""
    rng = np.random.default_rng(seed)
    s = 10.0 ** rng.uniform(-1, 1, A.rows)
    B = ol.CSR(A.rows, A.cols, A.rp, A.ci, A.v * np.repeat(s, np.diff(A.rp.astype(np.int64))))
    return B


def main():
    o = ol.oracle(); r = ol.ref(); r.set_threads(1)
    out = []
    for name, (nx, ny, nz), c, tol, maxit, m in CASES:
        A = vardiag(o.stencil3d(nx, ny, nz, *c) if nz > 1 else o.stencil2d(nx, ny, c[0], c[1]))
        b = np.ones(A.rows)
        res = r.solve("gmres", A, b, precond="jacobi", tol=tol, maxit=maxit, krylov=m)
        true = float(np.linalg.norm(b - A.to_scipy() @ res["x"]) / np.linalg.norm(b))
        out.append(dict(name=name, grid=[nx, ny, nz], c=list(c), tol=tol, maxit=maxit, krylov=m, iters=res["iters"], error=res["error"],
                        true_residual=true, x_norm=float(np.linalg.norm(res["x"]))))
        print(out[-1])
    json.dump(out, open(os.path.join(HERE, "gmres_jacobi.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
