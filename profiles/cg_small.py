"""Config 1 (CG, 2-D 5-point Laplacian 1024^2): iterations/s of the whole solve; under ncu: per-kernel durations."""
import sys, numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as ge
pkg = ge.load_package()
be = pkg.Backend(0)
maxit = int(sys.argv[1]) if len(sys.argv) > 1 else 1898
A = pkg.CsrMatrix.stencil(be, 1024, 1024, 1)
n = A.rows
b = be.array(np.ones(n)); x = be.zeros(n)
for rep in range(3):
    be.sync(); be.timer_begin()
    t = pkg.SolverTag(tol=1e-8, max_iterations=maxit).solve("cg", A, b, x)
    ms = be.timer_end()
    print("cg 1024^2: %d iters %.2f ms -> %.0f it/s (%.2f us/iter)" % (t.iters, ms, t.iters / ms * 1e3, ms * 1e3 / max(t.iters, 1)))
