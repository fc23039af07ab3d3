"""Tiny solves through every persistent kernel (CG, Jacobi-PCG, BiCGStab, GMRES with and without Jacobi) and the long-row / SELL-C-sigma
product paths -- meant to run under compute-sanitizer (memcheck / racecheck / synccheck), never timed."""
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import __graft_entry__ as ge
pkg = ge.load_package(); be = pkg.Backend(0)
A = pkg.CsrMatrix.stencil(be, 48, 50, 1, 0.5, 0.25, 0.0); L = pkg.CsrMatrix.stencil(be, 48, 50, 1)
n = A.rows
b = be.array(np.ones(n)); x = be.zeros(n)
print("cg", pkg.SolverTag(tol=1e-8, max_iterations=40).solve("cg", L, b, x).iters)
print("pcg", pkg.SolverTag(tol=1e-8, max_iterations=40, precond=1).solve("cg", L, b, x).iters)
print("bicgstab", pkg.SolverTag(tol=1e-8, max_iterations=20).solve("bicgstab", A, b, x).iters)
print("gmres", pkg.SolverTag(tol=1e-8, max_iterations=20, krylov_dim=10).solve("gmres", A, b, x).iters)
print("gmres+jacobi", pkg.SolverTag(tol=1e-8, max_iterations=20, krylov_dim=10, precond=1).solve("gmres", A, b, x).iters)
rng = np.random.default_rng(2)
rows = 700
lens = rng.integers(0, 12, rows); lens[5] = 300; lens[400] = 2500; lens[401] = 70
rp = np.zeros(rows + 1, np.uint32); rp[1:] = np.cumsum(lens)
ci = np.concatenate([np.sort(rng.choice(3000, l, replace=False)) for l in lens]).astype(np.uint32)
R = pkg.CsrMatrix.from_host(be, rows, 3000, rp, ci, rng.uniform(-1, 1, ci.size))
xr, yr = be.array(rng.uniform(1, 2, 3000)), be.zeros(rows)
R.spmv(xr, yr); R.to_sell(32).spmv(xr, yr); R.to_sell_sigma(32, 256).spmv(xr, yr)
print("products ok", float(np.abs(yr.download()).sum()) > 0)
be.close()
