import sys; sys.path.insert(0, ".")
import __graft_entry__ as ge
pkg = ge.load_package(); be = pkg.Backend(0)
A = pkg.CsrMatrix.stencil(be, 256, 256, 256)
n = A.rows
x, y = be.empty(n), be.zeros(n)
be.check(be.L.ViennaCLCUDADfill_uniform(be.h, n, x.ptr, 3, 0, 1.0, 2.0))
for name, M in (("ell", pkg.EllMatrix.from_csr(A)), ("hyb0.8", pkg.HybMatrix.from_csr(A, 0.8)), ("hyb0.01", pkg.HybMatrix.from_csr(A, 0.01))):
    for _ in range(5): M.spmv(x, y)
    be.sync(); be.timer_begin()
    for _ in range(20): M.spmv(x, y)
    ms = be.timer_end() / 20
    w = M.width if name == "ell" else M.ell.width
    nb = 12 * n * w + 16 * n
    print(name, "width", w, "ms", round(ms, 4), "GB/s (ELL bytes)", round(nb / ms / 1e6, 1))
b = be.array(__import__("numpy").ones(n))
import time
for name, M in (("csr", A), ("ell", pkg.EllMatrix.from_csr(A))):
    t = pkg.SolverTag(tol=0.0, max_iterations=100)
    t.solve("cg", M, b, y); be.sync()
    be.timer_begin(); t = pkg.SolverTag(tol=0.0, max_iterations=200); t.solve("cg", M, b, y); ms = be.timer_end()
    print("cg", name, round(200 / ms * 1e3, 1), "it/s")
