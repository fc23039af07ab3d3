"""Round 2: BASELINE config 1 (CG, 2-D 5-point Laplacian 1024^2) under the L2 policies of the matrix streams
(option "l2_resident": 0 evict-first as in round 1, 1 evict-last, 2 normal, -1 auto) and both driver forms
(persistent cooperative kernel / two kernels per iteration).  Prints iterations/s of whole converged solves."""
import sys
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as ge
pkg = ge.load_package()
be = pkg.Backend(0)
sizes = [int(a) for a in sys.argv[1:]] or [1024]
for n1 in sizes:
    A = pkg.CsrMatrix.stencil(be, n1, n1, 1)
    n = A.rows
    b = be.array(np.ones(n)); x = be.zeros(n)
    for persistent in (-1, 0):
        be.set_option("persistent_rows", persistent)
        for mode in (0, 1, 2, -1):
            be.set_option("l2_resident", mode)
            best = None
            for rep in range(3):
                be.sync(); be.timer_begin()
                t = pkg.SolverTag(tol=1e-8, max_iterations=5000).solve("cg", A, b, x)
                ms = be.timer_end()
                best = ms if best is None else min(best, ms)
            print("cg %d^2 %-11s l2_resident=%2d: %d iters %.2f ms -> %.0f it/s (%.2f us/iter)"
                  % (n1, "persistent" if persistent else "two-kernel", mode, t.iters, best, t.iters / best * 1e3, best * 1e3 / max(t.iters, 1)), flush=True)
    be.set_option("persistent_rows", -1); be.set_option("l2_resident", -1)
    # plain products on the same matrix, repeated: does the resident policy help a small matrix outside the solvers?
    y = be.zeros(n)
    for mode in (0, 1, 2):
        be.set_option("l2_resident", mode)
        for _ in range(5):
            A.spmv(b, y)
        be.sync(); be.timer_begin()
        for _ in range(200):
            A.spmv(b, y)
        ms = be.timer_end()
        print("spmv %d^2 l2_resident=%d: %.2f us per product (%.0f GB/s algorithmic)" % (n1, mode, ms * 1e3 / 200, A.bytes_spmv() * 200 / ms / 1e6), flush=True)
    be.set_option("l2_resident", -1)
