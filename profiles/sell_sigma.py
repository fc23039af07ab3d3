"""Irregular (power-law row lengths) matrix, 2M rows: padding and SpMV time of CSR, SELL-32 and SELL-32-sigma."""
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import __graft_entry__ as ge
pkg = ge.load_package(); be = pkg.Backend(0)
rng = np.random.default_rng(1)
rows = 2_000_000
lens = np.minimum((rng.pareto(1.5, rows) * 4).astype(np.int64) + 1, 2000)
rp = np.zeros(rows + 1, np.uint32); rp[1:] = np.cumsum(lens)
nnz = int(rp[-1])
ci = ((np.repeat(np.arange(rows), lens) + rng.integers(-50000, 50000, nnz)) % rows).astype(np.uint32)
va = rng.uniform(-1, 1, nnz)
A = pkg.CsrMatrix.from_host(be, rows, rows, rp, ci, va)
x, y = be.array(rng.uniform(1, 2, rows)), be.zeros(rows)
def timeit(M, name, nb):
    for _ in range(3): M.spmv(x, y)
    be.sync(); be.timer_begin()
    for _ in range(20): M.spmv(x, y)
    ms = be.timer_end() / 20
    print("%-16s stored entries %.1fM (x%.2f)  %.3f ms  %.0f GB/s of stored bytes, %.0f GB/s of CSR-equivalent bytes" %
          (name, nb / 1e6, nb / nnz, ms, (12 * nb + 16 * rows) / ms / 1e6, (12 * nnz + 20 * rows) / ms / 1e6), flush=True)
    return y.download()
y0 = timeit(A, "csr", nnz)
S1 = A.to_sell(32); y1 = timeit(S1, "sell-32", S1.padded_nnz)
for sg in (256, 1024, 4096):
    S = A.to_sell_sigma(32, sg); ys = timeit(S, "sell-32-%d" % sg, S.padded_nnz)
    assert np.array_equal(ys, y1)
print("max |csr - sell| = %.2e" % np.abs(y0 - y1).max())
