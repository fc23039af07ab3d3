import sys, numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as ge
pkg = ge.load_package(); be = pkg.Backend(0)
for (nx,ny,nz) in ((256,256,1),(1024,1024,1),(128,128,128)):
    A = pkg.CsrMatrix.stencil(be, nx, ny, nz); n=A.rows
    b = be.array(np.ones(n)); x = be.zeros(n)
    t=pkg.SolverTag(tol=1e-8, max_iterations=5000, precond=1).solve("cg", A, b, x)
    xs = x.download()
    pkg.SolverTag(tol=0.0, max_iterations=64, precond=1).solve("cg", A, b, x)
    be.sync(); be.timer_begin(); t2=pkg.SolverTag(tol=0.0, max_iterations=1000, precond=1).solve("cg", A, b, x); ms=be.timer_end()
    print((nx,ny,nz), "pcg iters", t.iters, "err %.2e" % t.error, "xnorm %.10e" % np.linalg.norm(xs), "%.1f us/iter" % (ms*1e3/t2.iters), flush=True)
