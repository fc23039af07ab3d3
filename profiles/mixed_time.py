"""256^3 Laplacian: plain double CG vs mixed-precision CG to the same tolerance (time to solution)."""
import sys, numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as ge
pkg = ge.load_package()
be = pkg.Backend(0)
A = pkg.CsrMatrix.stencil(be, 256, 256, 256)
n = A.rows
b = be.array(np.ones(n)); x = be.zeros(n)
for tol in (1e-6, 1e-8):
    be.sync(); be.timer_begin()
    t = pkg.SolverTag(tol=tol, max_iterations=5000).solve("cg", A, b, x)
    ms = be.timer_end()
    print("tol %g  double CG: %d iters, %.1f ms, error %.2e" % (tol, t.iters, ms, t.error))
    be.sync(); be.timer_begin()
    t = pkg.mixed_precision_cg(A, b, x, tol, 20000, 1e-2)
    ms = be.timer_end()
    print("tol %g  mixed  CG: %d iters, %.1f ms, error %.2e" % (tol, t.iters, ms, t.error))
