import sys, numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as ge
pkg = ge.load_package(); be = pkg.Backend(0)
for (nx,ny,nz) in ((2048,2048,1),(160,160,160),(200,200,200),(256,256,256)):
    A = pkg.CsrMatrix.stencil(be, nx, ny, nz); n=A.rows
    b = be.array(np.ones(n)); x = be.zeros(n)
    pkg.SolverTag(tol=0.0, max_iterations=64).solve("cg", A, b, x)
    be.sync(); be.timer_begin(); t2=pkg.SolverTag(tol=0.0, max_iterations=400).solve("cg", A, b, x); ms=be.timer_end()
    print((nx,ny,nz), "rows %.1fM" % (n/1e6), "%.1f us/iter" % (ms*1e3/t2.iters), "%.0f GB/s" % ((12*A.nnz+76*n)*t2.iters/ms/1e6), flush=True)
