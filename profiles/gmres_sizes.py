import sys, numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as ge
pkg = ge.load_package(); be = pkg.Backend(0)
for (nx,ny,nz,pc) in ((48,50,1,0),(48,50,1,1),(256,256,1,0),(512,512,1,0),(1024,1024,1,0),(128,128,128,0),(128,128,128,1)):
    c = (0.5,0.25,0.125) if nz > 1 else (0.5, 0.25, 0.0)
    A = pkg.CsrMatrix.stencil(be, nx, ny, nz, *c); n=A.rows
    b = be.array(np.ones(n)); x = be.zeros(n)
    t=pkg.SolverTag(tol=1e-8, max_iterations=3000, krylov_dim=30, precond=pc).solve("gmres", A, b, x)
    xs = x.download()
    pkg.SolverTag(tol=0.0, max_iterations=30, krylov_dim=30, precond=pc).solve("gmres", A, b, x)
    be.sync(); be.timer_begin(); t2=pkg.SolverTag(tol=0.0, max_iterations=300, krylov_dim=30, precond=pc).solve("gmres", A, b, x); ms=be.timer_end()
    print((nx,ny,nz), "precond", pc, "gmres(30) iters", t.iters, "err %.2e" % t.error, "xnorm %.10e" % np.linalg.norm(xs), "%.1f us/inner iter" % (ms*1e3/t2.iters), flush=True)
