"""CG 256^3 (multi-kernel driver), 200 iterations: the vector-update kernel compiled for 4 / 5 / 6 / 8 resident CTAs per SM."""
import sys
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as ge
pkg = ge.load_package()
be = pkg.Backend(0)
label = sys.argv[1] if len(sys.argv) > 1 else "head"
A = pkg.CsrMatrix.stencil(be, 256, 256, 256)
b = be.array(np.ones(A.rows)); x = be.zeros(A.rows)
be.set_option("persistent_rows", 0)
best = None
for rep in range(4):
    be.sync(); be.timer_begin()
    t = pkg.SolverTag(tol=0.0, max_iterations=200).solve("cg", A, b, x)
    ms = be.timer_end()
    best = ms if best is None else min(best, ms)
print("%-5s cg 256^3 200 iterations: %.2f ms -> %.0f it/s (%.1f us per iteration)" % (label, best, 200 / best * 1e3, best * 1e3 / 200), flush=True)
be.close()
