import sys, numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as ge
pkg = ge.load_package()
be = pkg.Backend(0)
for dt in (np.float32,):
    A = pkg.CsrMatrix.stencil(be, 256, 256, 256, dtype=dt)
    n = A.rows
    x, y = be.empty(n, dt), be.zeros(n, dt)
    be.check(be.lib_for(dt).ViennaCLCUDADfill_uniform(be.h, n, x.ptr, 3, 0, 1.0, 2.0))
    S = A.to_sell(32); E = pkg.EllMatrix.from_csr(A)
    for name, M in (("csr", A), ("sell", S), ("ell", E)):
        for _ in range(5): M.spmv(x, y)
        be.sync(); be.timer_begin()
        for _ in range(50): M.spmv(x, y)
        ms = be.timer_end() / 50
        esz = np.dtype(dt).itemsize
        byt = (esz + 4) * A.nnz + (4 if name == "csr" else 0) * n + 2 * esz * n
        print(np.dtype(dt).name, name, "ms %.4f" % ms, "GB/s %.0f" % (byt / ms / 1e6))
    b = be.array(np.ones(n, dt)); sol = be.zeros(n, dt)
    be.sync(); be.timer_begin()
    tag = pkg.SolverTag(tol=1e-30, max_iterations=200).solve("cg", A, b, sol)
    ms = be.timer_end()
    print("float cg 256^3: %d iters %.1f it/s" % (tag.iters, tag.iters / ms * 1e3))
