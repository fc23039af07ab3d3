"""512^3 CSR SpMV (BASELINE config 5 matrix on one GPU): time per product; under ncu: DRAM traffic vs the compulsory 13.94 GB."""
import sys, numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as ge
pkg = ge.load_package(); be = pkg.Backend(0)
n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 512
A = pkg.CsrMatrix.stencil(be, n1, n1, n1); n = A.rows
x, y = be.empty(n), be.zeros(n)
be.check(be.L.ViennaCLCUDADfill_uniform(be.h, n, x.ptr, 1, 0, 1.0, 2.0))
for _ in range(3): A.spmv(x, y)
be.sync(); be.timer_begin()
for _ in range(10): A.spmv(x, y)
ms = be.timer_end() / 10
byt = 12 * A.nnz + 20 * n
print("%d^3: %.3f ms per SpMV, %.0f GB/s, compulsory %.2f GB" % (n1, ms, byt / ms / 1e6, byt / 1e9))
