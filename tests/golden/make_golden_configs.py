"""Generates tests/golden/baseline_configs.json: the UNMODIFIED reference (OpenMP host backend, OMP_NUM_THREADS=1, compiled
by oracle/Makefile into oracle/_ref) run on the BASELINE.json configurations AT SIZE (or on the reduced grid SURVEY 8d
prescribes for C4) -- iteration counts, error estimates, true residuals, solution norms and a strided sample of x.

    python tests/golden/make_golden_configs.py [case ...]     (build container only: needs /root/reference; minutes per case)

Cases
  c1_cg_lap2d_1024             cg.hpp:128-187 pipelined CG, 2-D 5-point Laplacian 1024^2, b = 1, tol 1e-8
  c3_bicgstab_jacobi_cd3d_256  bicgstab.hpp:398-489 + jacobi_precond.hpp:103-130, 3-D upwind convection-diffusion 256^3
  c3_bicgstab_pipelined_cd3d_256  bicgstab.hpp:97-215 (no preconditioner), same system
  c4_gmres30_cd2d_512          gmres.hpp:449-631 (Householder path, identity preconditioner object) AND the pipelined path of the
                               gmres-fix build, 2-D upwind convection-diffusion 512^2, m = 30, tol 1e-8
  c4_gmres30_cd2d_4096_budget  the pipelined path of the gmres-fix build on the full 4096^2 grid, fixed budget of 60 iterations
  c2_cg_lap3d_256_budget       pipelined CG on the 3-D 7-point Laplacian 256^3, fixed budget of 20 iterations
  c5_cg_lap3d_512_budget       same on 512^3 (needs ~25 GB of host memory)
  c5_cg_lap3d_128              pipelined CG on 128^3 to convergence (the partition-independence anchor of SURVEY 8e)
Every case merges its record into the JSON file, so cases can be generated one at a time.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

OUT = os.path.join(HERE, "baseline_configs.json")
SAMPLE = 1024            # entries of x kept (evenly strided over the vector)


def sample_idx(n):
    """1024 entries spread over the whole vector by a multiplicative hash: a fixed stride would sit on one grid plane (a stride of
    16384 on the 256^3 grid only ever visits x = 0)."""
    return (np.arange(1, SAMPLE + 1, dtype=np.uint64) * np.uint64(2654435761) % np.uint64(n)).astype(np.int64)


def true_residual(o, A, x, b):
    return float(np.linalg.norm(b - o.csr_spmv(A, x)) / np.linalg.norm(b))


def record(name, A, b, res, o, extra):
    idx = sample_idx(A.rows)
    rec = dict(name=name, rows=A.rows, nnz=A.nnz, iters=int(res["iters"]), error=float(res["error"]),
               true_residual=true_residual(o, A, res["x"], b), x_norm=float(np.linalg.norm(res["x"])),
               x_sample_rule="idx_k = (k * 2654435761) mod rows, k = 1..1024", x_sample=[float(v) for v in res["x"][idx]], seconds=float(res["seconds"]), threads=1)
    rec.update(extra)
    return rec


def run_case(name):
    o = ol.oracle()
    r = ol.ref()
    rf = ol.ref(fixed=True)
    # omp_set_num_threads is process-wide (both checkers share libgomp): every solve below runs on ONE thread
    r.set_threads(1); rf.set_threads(1); o.set_threads(1)
    t0 = time.time()
    if name == "c1_cg_lap2d_1024":
        A = o.stencil2d(1024, 1024); b = np.ones(A.rows)
        res = r.solve("cg", A, b, tol=1e-8, maxit=5000)
        rec = record(name, A, b, res, o, dict(solver="cg", tol=1e-8, maxit=5000, grid=[1024, 1024, 1], c=[0, 0, 0]))
    elif name in ("c3_bicgstab_jacobi_cd3d_256", "c3_bicgstab_pipelined_cd3d_256"):
        A = o.stencil3d(256, 256, 256, 0.5, 0.25, 0.125); b = np.ones(A.rows)
        pre = "jacobi" if "jacobi" in name else "none"
        res = r.solve("bicgstab", A, b, precond=pre, tol=1e-8, maxit=2000)
        rec = record(name, A, b, res, o, dict(solver="bicgstab", precond=pre, tol=1e-8, maxit=2000, grid=[256, 256, 256], c=[0.5, 0.25, 0.125]))
    elif name == "c4_gmres30_cd2d_512":
        A = o.stencil2d(512, 512, 0.5, 0.25); b = np.ones(A.rows)
        res = r.solve("gmres", A, b, precond="identity", tol=1e-8, maxit=30000, krylov=30)
        rec = record(name, A, b, res, o, dict(solver="gmres", path="householder (gmres.hpp:449-631)", tol=1e-8, maxit=30000, krylov=30,
                                              grid=[512, 512, 1], c=[0.5, 0.25, 0]))
        res2 = rf.solve("gmres", A, b, precond="none", tol=1e-8, maxit=30000, krylov=30)
        rec["pipelined_gmresfix"] = dict(iters=int(res2["iters"]), error=float(res2["error"]), true_residual=true_residual(o, A, res2["x"], b),
                                         x_norm=float(np.linalg.norm(res2["x"])), x_sample=[float(v) for v in res2["x"][sample_idx(A.rows)]])
    elif name == "c4_gmres30_cd2d_4096_budget":
        A = o.stencil2d(4096, 4096, 0.5, 0.25); b = np.ones(A.rows)
        res = rf.solve("gmres", A, b, precond="none", tol=1e-10, maxit=60, krylov=30)
        rec = record(name, A, b, res, o, dict(solver="gmres", path="pipelined, gmres-fix build", tol=1e-10, maxit=60, krylov=30,
                                              grid=[4096, 4096, 1], c=[0.5, 0.25, 0]))
    elif name in ("c2_cg_lap3d_256_budget", "c5_cg_lap3d_512_budget"):
        n1 = 256 if "256" in name else 512
        A = o.stencil3d(n1, n1, n1); b = np.ones(A.rows)
        res = r.solve("cg", A, b, tol=1e-30, maxit=20)
        rec = record(name, A, b, res, o, dict(solver="cg", tol=1e-30, maxit=20, grid=[n1, n1, n1], c=[0, 0, 0]))
    elif name == "c5_cg_lap3d_128":
        A = o.stencil3d(128, 128, 128); b = np.ones(A.rows)
        res = r.solve("cg", A, b, tol=1e-8, maxit=2000)
        rec = record(name, A, b, res, o, dict(solver="cg", tol=1e-8, maxit=2000, grid=[128, 128, 128], c=[0, 0, 0]))
    else:
        raise SystemExit("unknown case " + name)
    rec["wall_seconds"] = time.time() - t0
    return rec


ALL = ["c1_cg_lap2d_1024", "c3_bicgstab_jacobi_cd3d_256", "c3_bicgstab_pipelined_cd3d_256", "c4_gmres30_cd2d_512",
       "c4_gmres30_cd2d_4096_budget", "c2_cg_lap3d_256_budget", "c5_cg_lap3d_512_budget", "c5_cg_lap3d_128"]


def main():
    cases = sys.argv[1:] or ALL
    for name in cases:
        rec = run_case(name)
        print({k: v for k, v in rec.items() if k != "x_sample" and k != "pipelined_gmresfix"}, flush=True)
        # merge under a lock-free read-modify-write (cases may be generated by parallel processes: re-read just before writing)
        part = OUT + "." + name + ".part"
        json.dump(rec, open(part, "w"))
    merged = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for f in sorted(os.listdir(HERE)):
        if f.startswith("baseline_configs.json.") and f.endswith(".part"):
            rec = json.load(open(os.path.join(HERE, f)))
            merged[rec["name"]] = rec
            os.remove(os.path.join(HERE, f))
    json.dump(merged, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
