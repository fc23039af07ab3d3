"""Launches every hot kernel a few times at benchmark size so that ncu can capture it (never a bench number).

    ncu --set full --clock-control none --import-source on -k regex:<name> -s 1 -c 2 -o gpurun_out/prof_<name> \
        python profiles/run_kernels.py <what>

what: spmv (CSR + SELL, 256^3) | spmv_split (the SPLIT instantiation, world 1) | cg (3 iterations at 256^3, CSR) | cg_sell | bicgstab | bicgstab_jacobi | gmres | gmres_jacobi |
      cg1024 (persistent cooperative kernel) | spmv_f32 (float CSR / SELL / ELL, 256^3)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
what = sys.argv[1] if len(sys.argv) > 1 else "spmv"
be = pkg.Backend(0)
n1 = 256

if what == "cg1024":
    A = pkg.CsrMatrix.stencil(be, 1024, 1024, 1)
else:
    c = (0.5, 0.25, 0.125) if what.startswith("bicgstab") or what.startswith("gmres") else (0.0, 0.0, 0.0)
    A = pkg.CsrMatrix.stencil(be, n1, n1, n1, *c)
n = A.rows
x, y = be.empty(n), be.zeros(n)
be.check(be.L.ViennaCLCUDADfill_uniform(be.h, n, x.ptr, 1, 0, 1.0, 2.0))
b = be.array(np.ones(n))

if what == "spmv_f32":
    F = np.float32
    Af = pkg.CsrMatrix.stencil(be, n1, n1, n1, dtype=F)
    xf, yf = be.empty(n, F), be.zeros(n, F)
    be.check(be.lib_for(F).ViennaCLCUDADfill_uniform(be.h, n, xf.ptr, 1, 0, 1.0, 2.0))
    Sf = Af.to_sell(32); Ef = pkg.EllMatrix.from_csr(Af)
    for _ in range(3):
        Af.spmv(xf, yf); Sf.spmv(xf, yf); Ef.spmv(xf, yf)
elif what == "gmres_jacobi":
    pkg.SolverTag(tol=0.0, max_iterations=30, krylov_dim=30, precond=1).solve("gmres", A, b, y)
elif what == "spmv_split":
    # the row-partitioned instantiation csr_stream_kernel<EpiAxpby, SPLIT=true> on one rank (no halo: all row blocks are interior)
    D = pkg.DistCsr(be, n, 0, n, A)
    for _ in range(3):
        D.spmv(x, y)
elif what == "spmv":
    S = A.to_sell(32)
    for _ in range(3):
        A.spmv(x, y)
        S.spmv(x, y)
elif what in ("cg", "cg1024"):
    if len(sys.argv) > 2:
        be.set_option("persistent_cg_form", int(sys.argv[2]))
    pkg.SolverTag(tol=0.0, max_iterations=32 if what == "cg1024" else 4).solve("cg", A, b, y)
elif what == "cg_sell":
    pkg.SolverTag(tol=0.0, max_iterations=4).solve("cg", A.to_sell(32), b, y)
elif what == "bicgstab":
    pkg.SolverTag(tol=0.0, max_iterations=3).solve("bicgstab", A, b, y)
elif what == "bicgstab_jacobi":
    pkg.SolverTag(tol=0.0, max_iterations=3, precond=1).solve("bicgstab", A, b, y)
elif what == "gmres":
    pkg.SolverTag(tol=0.0, max_iterations=30, krylov_dim=30).solve("gmres", A, b, y)
be.sync()
print("done", what, be.launches(), "launches")
be.close()
