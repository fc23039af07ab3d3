"""Per-iteration time of the fused solver paths at 256^3 (fixed iteration budgets): Jacobi-BiCGStab, BiCGStab, GMRES(30) with and
without Jacobi, CG on CSR / SELL / ELL.  Used to compare kernel variants (VCL_B200_LIB_OVERRIDE)."""
import sys, numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as ge
pkg = ge.load_package(); be = pkg.Backend(0)
A = pkg.CsrMatrix.stencil(be, 256, 256, 256, 0.5, 0.25, 0.125); n = A.rows
b = be.array(np.ones(n)); x = be.zeros(n)
def run(name, solver, M, iters, **kw):
    pkg.SolverTag(tol=0.0, max_iterations=min(iters, 30), **kw).solve(solver, M, b, x)
    be.sync(); be.timer_begin(); t = pkg.SolverTag(tol=0.0, max_iterations=iters, **kw).solve(solver, M, b, x); ms = be.timer_end()
    print("%-28s %4d iters  %.1f us/iter" % (name, t.iters, ms * 1e3 / max(t.iters, 1)), flush=True)
run("bicgstab+jacobi csr", "bicgstab", A, 100, precond=1)
run("bicgstab csr", "bicgstab", A, 100)
run("gmres(30) csr", "gmres", A, 90, krylov_dim=30)
run("gmres(30)+jacobi csr", "gmres", A, 90, krylov_dim=30, precond=1)
L = pkg.CsrMatrix.stencil(be, 256, 256, 256)
run("cg csr", "cg", L, 100)
run("cg+jacobi csr", "cg", L, 100, precond=1)
run("cg sell32", "cg", L.to_sell(32), 100)
E_ = pkg.EllMatrix.from_csr(L)
run("cg ell", "cg", E_, 100)
run("bicgstab ell", "bicgstab", pkg.EllMatrix.from_csr(A), 100)
