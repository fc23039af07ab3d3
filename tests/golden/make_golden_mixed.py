"""Generates tests/golden/mixed_precision_cg.json: iteration counts / errors of the UNMODIFIED reference's
mixed_precision_cg (linalg/mixed_precision_cg.hpp, host backend, 1 thread) on small systems.

    python tests/golden/make_golden_mixed.py      (build container only: needs /root/reference)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

CASES = [("lap2d_63x65", (63, 65, 1), 1e-8, 1000, 1e-2), ("lap2d_63x65", (63, 65, 1), 1e-10, 3000, 1e-3),
         ("lap3d_20", (20, 20, 20), 1e-8, 1000, 1e-2), ("lap2d_200", (200, 200, 1), 1e-8, 2000, 1e-2),
         ("lap2d_200", (200, 200, 1), 1e-8, 50, 1e-2)]


def main():
    o = ol.oracle(); r = ol.ref(); r.set_threads(1)
    out = []
    for name, (nx, ny, nz), tol, maxit, itol in CASES:
        A = o.stencil3d(nx, ny, nz) if nz > 1 else o.stencil2d(nx, ny)
        b = np.ones(A.rows)
        res = r.mixed_cg(A, b, tol, maxit, itol)
        plain = r.solve("cg", A, b, tol=tol, maxit=maxit)
        true = float(np.linalg.norm(b - A.to_scipy() @ res["x"]) / np.linalg.norm(b))
        out.append(dict(name=name, grid=[nx, ny, nz], tol=tol, maxit=maxit, inner_tol=itol, iters=res["iters"], error=res["error"],
                        true_residual=true, plain_cg_iters=plain["iters"], x_norm=float(np.linalg.norm(res["x"]))))
        print(out[-1])
    json.dump(out, open(os.path.join(HERE, "mixed_precision_cg.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
