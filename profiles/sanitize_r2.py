"""Round-2 kernels under compute-sanitizer (memcheck / racecheck), never timed: both persistent CG forms (one-pass, two-phase) on
small systems, the foreign-plan check and the plan-free fallback, and the row-partitioned drivers at world 1 (SPLIT instantiation,
CG / Jacobi-CG / BiCGStab / GMRES over slabs)."""
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import __graft_entry__ as ge
pkg = ge.load_package(); be = pkg.Backend(0)
L = pkg.CsrMatrix.stencil(be, 48, 50, 1); A = pkg.CsrMatrix.stencil(be, 17, 13, 11, 0.5, 0.25, 0.125)
n = L.rows
b = be.array(np.ones(n)); x = be.zeros(n)
for form in (1, 3, 2):
    be.set_option("persistent_cg_form", form)
    print("cg form", form, pkg.SolverTag(tol=1e-8, max_iterations=70).solve("cg", L, b, x).iters)
be.set_option("persistent_cg_form", 0)
be.set_option("persistent_rows", 0)
print("cg two-kernel", pkg.SolverTag(tol=1e-8, max_iterations=40).solve("cg", L, b, x).iters)
be.set_option("persistent_rows", -1)
# foreign plans: valid (copy of the own plan), breaking the limits (one block), small blocks
y = be.zeros(n)
own = L.blocks.download()[:L.nblocks + 1]
for name, blk in (("own_copy", own.copy()), ("one_block", np.array([0, n], np.uint32)), ("pairs", np.arange(0, n + 1, 2, dtype=np.uint32))):
    L.blocks = be.array(blk); L.nblocks = len(blk) - 1
    L.spmv(b, y)
    print("plan", name, float(y.download().sum()))
# row-partitioned drivers, world 1
na = A.rows
D = pkg.DistCsr(be, na, 0, na, A)
ba = be.array(np.ones(na)); xa = be.zeros(na); ya = be.zeros(na)
D.spmv(ba, ya)
print("dist cg", D.cg(ba, xa, pkg.SolverTag(tol=1e-8, max_iterations=30)).iters)
print("dist pcg", D.cg(ba, xa, pkg.SolverTag(tol=1e-8, max_iterations=30, precond=1)).iters)
print("dist bicgstab", D.bicgstab(ba, xa, pkg.SolverTag(tol=1e-8, max_iterations=30)).iters)
print("dist gmres", D.gmres(ba, xa, pkg.SolverTag(tol=1e-8, max_iterations=20, krylov_dim=10)).iters)
D.close()
be.close()
print("done")
