"""The SPLIT instantiations (row-partitioned slabs) at world 1 against the plain kernels, 256^3: CSR and SELL-32.  Kernel experiments."""
import sys
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as ge
pkg = ge.load_package()
be = pkg.Backend(0)
A = pkg.CsrMatrix.stencil(be, 256, 256, 256)
S = A.to_sell(32)
n = A.rows
x, y = be.empty(n), be.zeros(n)
be.check(be.L.ViennaCLCUDADfill_uniform(be.h, n, x.ptr, 1, 0, 1.0, 2.0))
D = pkg.DistCsr(be, n, 0, n, A)
def t(f):
    best = None
    for rep in range(3):
        for _ in range(5): f()
        be.sync(); be.timer_begin()
        for _ in range(50): f()
        ms = be.timer_end() / 50
        best = ms if best is None else min(best, ms)
    return best
print("csr  plain  %.4f ms" % t(lambda: A.spmv(x, y)))
print("sell plain  %.4f ms" % t(lambda: S.spmv(x, y)))
print("csr  split  %.4f ms" % t(lambda: D.spmv(x, y)))
D.set_format("sell", 32)
print("sell split  %.4f ms" % t(lambda: D.spmv(x, y)))
be.close()
