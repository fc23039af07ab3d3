"""Generates tests/golden/bicgstab_spread.json: how far the UNMODIFIED reference's own BiCGStab iteration counts move when only
the ORDER of its inner-product summation changes (OpenMP thread count 1 / 2 / 4 / 8 -> different partial-sum grouping,
host_based/vector_operations.hpp:463-472), on BASELINE config 3 (3-D upwind convection-diffusion 256^3, b = 1, tol 1e-8).
BiCGStab is not backward stable with respect to such perturbations; this file is the measured yardstick the
GPU tests use for the pipelined variant instead of a fixed +-2.

    python tests/golden/make_golden_spread.py        (build container only; ~15 minutes)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402


def small(out):
    """The small solver goldens of make_golden.py (BiCGStab, no preconditioner / Jacobi) at 1 / 2 / 3 / 4 / 8 threads."""
    o = ol.oracle(); r = ol.ref()
    for name, A in (("lap2d_63x65", o.stencil2d(63, 65)), ("cd2d_48x50", o.stencil2d(48, 50, 0.5, 0.0)), ("cd3d_11x10x9", o.stencil3d(11, 10, 9, 0.5, 0.25, 0.125))):
        b = np.ones(A.rows)
        for pre in ("none", "jacobi"):
            key = "small/%s/bicgstab_%s" % (name, pre)
            out[key] = {}
            for t in (1, 2, 3, 4, 8):
                r.set_threads(t)
                res = r.solve("bicgstab", A, b, precond=pre, tol=1e-8, maxit=1000)
                out[key][str(t)] = dict(iters=int(res["iters"]), error=float(res["error"]))
            print(key, {t: v["iters"] for t, v in out[key].items()}, flush=True)
            if pre == "jacobi":
                out[key + "/summation_orders"] = generic_bicgstab_orders(A, pre)
                print(key + "/summation_orders", {t: v["iters"] for t, v in out[key + "/summation_orders"].items()}, flush=True)


def generic_bicgstab_orders(A, precond):
    """The reference's generic (preconditioned) BiCGStab, bicgstab.hpp:398-489, restated in numpy with the inner products summed in
    different ORDERS -- sequential (the reference at one thread), exactly rounded (math.fsum), pairwise (np.sum), BLAS (np.dot) and
    k interleaved partial sums.  Everything else is identical.  Shows how far the iteration count moves under admissible
    summation orders: the yardstick for a device whose reductions are necessarily grouped differently."""
    import math
    S = A.to_scipy(); n = A.rows
    b = np.ones(n); diag = S.diagonal() if precond == "jacobi" else np.ones(n)

    def seq(a, c):
        t = 0.0
        for v in (a * c):
            t += v
        return t

    orders = {"sequential": seq, "exact_fsum": lambda a, c: math.fsum(a * c), "pairwise_np_sum": lambda a, c: float(np.sum(a * c)), "blas_np_dot": lambda a, c: float(np.dot(a, c))}
    for k in (2, 3, 7, 32, 256):
        orders["interleaved_%d" % k] = (lambda k: lambda a, c: float(sum(np.sum((a * c)[j::k]) for j in range(k))))(k)
    out = {}
    for name, dot in orders.items():
        x = np.zeros(n); r = (b - S @ x) / diag; p = r.copy(); r0 = r.copy()
        nb = math.sqrt(dot(b, b)); ip = math.sqrt(dot(r, r)); ip *= ip
        its = -1
        for i in range(1000):
            t0 = (S @ p) / diag
            alpha = ip / dot(t0, r0)
            s = r - alpha * t0
            t1 = (S @ s) / diag
            nt = math.sqrt(dot(t1, t1))
            omega = dot(t1, s) / (nt * nt)
            x += alpha * p + omega * s
            r = s - omega * t1
            if math.sqrt(dot(r, r)) / nb < 1e-8:
                its = i + 1
                break
            new = dot(r, r0)
            beta = new / ip * alpha / omega
            ip = new
            p -= omega * t0
            p = r + beta * p
        out[name] = dict(iters=its)
    return out


def main():
    o = ol.oracle(); r = ol.ref()
    path = os.path.join(HERE, "bicgstab_spread.json")
    if "--small" in sys.argv:                   # add / refresh only the small cases (seconds)
        out = json.load(open(path))
        small(out)
        json.dump(out, open(path, "w"), indent=1, sort_keys=True)
        return
    out = {}
    small(out)
    for grid, c in (((256, 256, 256), (0.5, 0.25, 0.125)), ((128, 128, 128), (0.5, 0.25, 0.125))):
        o.set_threads(8)
        A = o.stencil3d(*grid, *c); b = np.ones(A.rows)
        for pre in ("none", "jacobi"):
            key = "cd3d_%d_%s" % (grid[0], "pipelined" if pre == "none" else "jacobi")
            out[key] = {}
            for t in (1, 2, 4, 8):
                r.set_threads(t)
                res = r.solve("bicgstab", A, b, precond=pre, tol=1e-8, maxit=2000, hist_cap=2000)
                o.set_threads(8)
                true = float(np.linalg.norm(b - o.csr_spmv(A, res["x"])) / np.linalg.norm(b))
                out[key][str(t)] = dict(iters=int(res["iters"]), error=float(res["error"]), true_residual=true,
                                        history_every_25=[float(v) for v in res["history"][::25]])
                print(key, t, out[key][str(t)]["iters"], res["error"], true, flush=True)
                json.dump(out, open(os.path.join(HERE, "bicgstab_spread.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
