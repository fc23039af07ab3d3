"""Stage-depth experiment (library variants via VCL_B200_LIB_OVERRIDE): CSR product on 2-D 5-point and 3-D 7-point grids and the
config-1 CG through the three driver forms.  Kernel experiments; never a bench number."""
import sys
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as ge
pkg = ge.load_package()
be = pkg.Backend(0)
label = sys.argv[1] if len(sys.argv) > 1 else "head"
for shape in ((1024, 1024, 1), (2048, 2048, 1), (4096, 4096, 1), (128, 128, 128), (256, 256, 256)):
    A = pkg.CsrMatrix.stencil(be, *shape)
    n = A.rows
    x, y = be.empty(n), be.zeros(n)
    be.check(be.L.ViennaCLCUDADfill_uniform(be.h, n, x.ptr, 1, 0, 1.0, 2.0))
    best = None
    for rep in range(3):
        for _ in range(5):
            A.spmv(x, y)
        be.sync(); be.timer_begin()
        for _ in range(100):
            A.spmv(x, y)
        ms = be.timer_end() / 100
        best = ms if best is None else min(best, ms)
    print("%-6s spmv %-14s %8.2f us  %6.0f GB/s" % (label, "x".join(map(str, shape)), best * 1e3, A.bytes_spmv() / best / 1e6), flush=True)
    del A, x, y
for n1 in (1024,):
    A = pkg.CsrMatrix.stencil(be, n1, n1, 1)
    n = A.rows
    b = be.array(np.ones(n)); x = be.zeros(n)
    for name, rows, form in (("onepass", -1, 1), ("twophase", -1, 2), ("two-kernel", 0, 1)):
        be.set_option("persistent_rows", rows); be.set_option("persistent_cg_form", form)
        best = None
        for rep in range(3):
            be.sync(); be.timer_begin()
            t = pkg.SolverTag(tol=1e-8, max_iterations=5000).solve("cg", A, b, x)
            ms = be.timer_end()
            best = ms if best is None else min(best, ms)
        print("%-6s cg %d^2 %-11s: %d iters %.2f ms -> %.0f it/s (%.2f us/iter)" % (label, n1, name, t.iters, best, t.iters / best * 1e3, best * 1e3 / max(t.iters, 1)), flush=True)
A = pkg.CsrMatrix.stencil(be, 256, 256, 256)
b = be.array(np.ones(A.rows)); x = be.zeros(A.rows)
be.set_option("persistent_rows", 0)
for rep in range(2):
    be.sync(); be.timer_begin()
    t = pkg.SolverTag(tol=0.0, max_iterations=100).solve("cg", A, b, x)
    ms = be.timer_end()
print("%-6s cg 256^3 two-kernel 100 iterations: %.2f ms -> %.0f it/s" % (label, ms, 100 / ms * 1e3), flush=True)
be.close()
