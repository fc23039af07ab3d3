"""A/B timing of the 256^3 CSR / SELL SpMV over library variants (kernel experiments; never a bench number).
    python profiles/ab_spmv.py <tree root> [label]      -- run once per variant, each in its own process"""
import os
import sys
root = os.path.abspath(sys.argv[1]); label = sys.argv[2] if len(sys.argv) > 2 else root
sys.path.insert(0, root)
import __graft_entry__ as ge
pkg = ge.load_package()
be = pkg.Backend(0)
n1 = 256
A = pkg.CsrMatrix.stencil(be, n1, n1, n1)
S = A.to_sell(32)
n = A.rows
x, y = be.empty(n), be.zeros(n)
be.check(be.L.ViennaCLCUDADfill_uniform(be.h, n, x.ptr, 1, 0, 1.0, 2.0))
for name, M in (("csr", A), ("sell", S)):
    best = None
    for rep in range(5):
        for _ in range(5):
            M.spmv(x, y)
        be.sync(); be.timer_begin()
        for _ in range(50):
            M.spmv(x, y)
        ms = be.timer_end() / 50
        best = ms if best is None else min(best, ms)
    print("%-28s %-4s %.4f ms  %.0f GB/s" % (label, name, best, M.bytes_spmv() / best / 1e6), flush=True)
be.close()
