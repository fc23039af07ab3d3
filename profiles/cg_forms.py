"""Round 2: BASELINE config 1 (CG, 2-D 5-point Laplacian n^2) through the three driver forms:
one-pass persistent kernel (option persistent_cg_form 1: 2 CTAs/SM, 3: 3 CTAs/SM), two-phase persistent kernel (2), two kernels
per iteration (persistent_rows 0).  Whole converged solves, best of 3, iterations/s."""
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
    for name, rows, form in (("onepass/2cta", -1, 1), ("onepass/3cta", -1, 3), ("twophase", -1, 2), ("two-kernel", 0, 1)):
        be.set_option("persistent_rows", rows); be.set_option("persistent_cg_form", form)
        best = None
        for rep in range(3):
            be.sync(); be.timer_begin()
            t = pkg.SolverTag(tol=1e-8, max_iterations=5000).solve("cg", A, b, x)
            ms = be.timer_end()
            best = ms if best is None else min(best, ms)
        print("cg %d^2 %-13s: %d iters err %.6e  %.2f ms -> %.0f it/s (%.2f us/iter)"
              % (n1, name, t.iters, t.error, best, t.iters / best * 1e3, best * 1e3 / max(t.iters, 1)), flush=True)
    be.set_option("persistent_rows", -1); be.set_option("persistent_cg_form", 0)
