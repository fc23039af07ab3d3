"""Row-partitioned SpMV / CG parity check, launched one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

Every rank builds the same global matrix with the CPU oracle, keeps its slab of rows (global column indices), and compares
the distributed product with the oracle's product bit for bit and the distributed CG with the oracle's single-domain CG
(iteration count independent of the number of GPUs within +-2).  Also exercised with world = 1."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import oracle_lib as ol  # noqa: E402


def slab(A, rb, re_):
    k0, k1 = int(A.rp[rb]), int(A.rp[re_])
    return (A.rp[rb:re_ + 1] - A.rp[rb]).astype(np.uint32), A.ci[k0:k1], A.v[k0:k1]


def random_spd(n, seed):
    """Symmetric, diagonally dominant, with long-range couplings (halo from non-neighbouring ranks, ragged rows)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    m = 6 * n
    i, j = rng.integers(0, n, m), rng.integers(0, n, m)
    v = rng.uniform(-1, 1, m)
    B = sp.coo_matrix((v, (i, j)), shape=(n, n)).tocsr()
    S = B + B.T
    S.setdiag(0)
    S.eliminate_zeros()
    d = np.asarray(abs(S).sum(axis=1)).ravel() + 1.0
    S = (S + sp.diags(d)).tocsr()
    S.sort_indices()
    return ol.CSR(n, n, S.indptr.astype(np.uint32), S.indices.astype(np.uint32), S.data)


def big_case(pkg, be, rank, world):
    """CG on the 128^3 Laplacian to 1e-8 on every partitioning against the UNMODIFIED reference's single-domain run
    (tests/golden/baseline_configs.json, case c5_cg_lap3d_128: count +-2, same estimate, same x at the sampled entries)."""
    import json
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "baseline_configs.json")))["c5_cg_lap3d_128"]
    nx, ny, nz = g["grid"]
    n = nx * ny * nz
    rb, re_ = n * rank // world, n * (rank + 1) // world
    dA = pkg.CsrMatrix.stencil(be, nx, ny, nz, row_begin=rb, row_end=re_)
    D = pkg.DistCsr(be, n, rb, re_, dA)
    db, dx = be.array(np.ones(re_ - rb)), be.zeros(re_ - rb)
    tag = D.cg(db, dx, pkg.SolverTag(tol=g["tol"], max_iterations=g["maxit"]))
    idx = (np.arange(1, 1025, dtype=np.uint64) * np.uint64(2654435761) % np.uint64(n)).astype(np.int64)
    mine = (idx >= rb) & (idx < re_)
    x = dx.download()
    xs, xr = x[idx[mine] - rb], np.asarray(g["x_sample"])[mine]
    d = float(np.abs(xs - xr).max() / np.abs(np.asarray(g["x_sample"])).max()) if mine.any() else 0.0
    good = abs(tag.iters - g["iters"]) <= 2 and tag.error < g["tol"] and d <= 1e-6
    print("[rank %d/%d] lap3d_128^3 (reference golden): cg iters %d (reference %d) err %.3e (reference %.3e) max sample diff %.2e  %s [%s]"
          % (rank, world, tag.iters, g["iters"], tag.error, g["error"], d, "OK" if good else "FAIL", D.info()["transport"]), flush=True)
    D.close()
    return good


def main():
    import torch
    import torch.distributed as dist
    pkg = ge.load_package()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    be = pkg.Backend(local)
    if world > 1:
        ids = [be.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        be.comm_init(ids[0], rank, world)
    o = ol.oracle()
    o.set_threads(4)
    ok = True

    cases = [("lap3d_24x20x%d" % (8 * world), o.stencil3d(24, 20, 8 * world)),
             ("cd3d_17x13x%d" % (5 * world + 1), o.stencil3d(17, 13, 5 * world + 1, 0.5, 0.25, 0.125)),
             ("random_spd_5000", random_spd(5000, 3)),
             ("lap2d_300x301", o.stencil2d(300, 301))]
    for name, A in cases:
        n = A.rows
        bounds = [n * r // world for r in range(world + 1)]
        if name.startswith("random"):                       # uneven slabs
            bounds = [0] + [min(n, n * (r + 1) // world + (37 if r % 2 == 0 else -11)) for r in range(world - 1)] + [n]
        rb, re_ = bounds[rank], bounds[rank + 1]
        rp, ci, va = slab(A, rb, re_)
        dA = pkg.CsrMatrix.from_host(be, re_ - rb, n, rp, ci, va, with_blocks=False)
        D = pkg.DistCsr(be, n, rb, re_, dA)
        info = D.info()
        if world > 1 and os.environ.get("VCL_B200_DIST_TRANSPORT", "p2p") != "nccl" and os.environ.get("VCL_EXPECT_P2P", "1") == "1":
            ok &= info["transport"] == "peer-memory"      # NVLink box: the IPC windows must map
        x = o.uniform(n, 11, 1.0, 2.0)
        y_ref = o.csr_spmv(A, x)
        dx, dy = be.array(x[rb:re_]), be.zeros(re_ - rb)
        for _ in range(3):
            D.spmv(dx, dy)
        y = dy.download()
        same = np.array_equal(y, y_ref[rb:re_])
        b = np.ones(n)
        ref = o.cg(A, b, tol=1e-9, maxit=2000)
        db, dsol = be.array(b[rb:re_]), be.zeros(re_ - rb)
        tag = D.cg(db, dsol, pkg.SolverTag(tol=1e-9, max_iterations=2000))
        sol = dsol.download()
        err = np.linalg.norm(sol - ref["x"][rb:re_]) / max(np.linalg.norm(ref["x"][rb:re_]), 1e-300)
        good = same and abs(tag.iters - ref["iters"]) <= 2 and err < 1e-6 and (tag.error < 1e-9 or ref["error"] >= 1e-9)   # (CG on the nonsymmetric case stagnates, identically)
        ok &= good
        print("[rank %d/%d] %-22s spmv bit-exact=%s  cg iters %d (oracle %d) err %.2e rel-x-diff %.2e  %s"
              % (rank, world, name, same, tag.iters, ref["iters"], tag.error, err, "OK" if good else "FAIL") + " [%s]" % info["transport"], flush=True)
        # ---- row-partitioned BiCGStab (pipelined) and Jacobi-CG against the single-domain library on the full matrix (every rank
        # holds the whole small matrix too) and against the oracle's count; the single-domain library is itself pinned to the
        # reference by tests/test_gpu_parity.py / test_gpu_configs.py ----
        full = pkg.CsrMatrix.from_host(be, n, n, A.rp, A.ci, A.v)
        dbf = be.array(b)
        for solver in ("bicgstab", "pcg", "gmres"):
            if solver == "pcg" and name.startswith("cd"):
                continue                                     # CG on the nonsymmetric case: covered above (stagnates identically)
            kw = dict(tol=1e-9, max_iterations=2000, precond=(1 if solver == "pcg" else 0))
            if solver == "gmres":                            # GMRES(20), at most 30 cycles (the 2-D Laplacian exhausts the budget: same iterate required)
                kw = dict(tol=1e-9, max_iterations=600, krylov_dim=20)
            dxf = be.zeros(n)
            t1 = pkg.SolverTag(**kw).solve({"pcg": "cg"}.get(solver, solver), full, dbf, dxf)
            dsol2 = be.zeros(re_ - rb)
            tg = pkg.SolverTag(**kw)
            t2 = D.bicgstab(db, dsol2, tg) if solver == "bicgstab" else (D.gmres(db, dsol2, tg) if solver == "gmres" else D.cg(db, dsol2, tg))
            xs, xf = dsol2.download(), dxf.download()[rb:re_]
            err2 = np.linalg.norm(xs - xf) / max(np.linalg.norm(xf), 1e-300)
            o_it = o.bicgstab(A, b, tol=1e-9, maxit=2000)["iters"] if solver == "bicgstab" else None
            # BiCGStab counts move with the grouping of the inner-product sums (tests/golden/bicgstab_spread.json); the partitioned
            # sums are grouped per rank (lap2d_300x301: 500 on two ranks, 537 on one, the oracle between 521 and 545 from run to run at 4
            # threads).  Bar: same tolerance reached, same solution, count within 2 (PCG) / 10 % + 2 (BiCGStab).
            slack = 2 + t1.iters // 10 if solver == "bicgstab" else 2
            reached = t2.error < 1e-9 or (solver == "gmres" and t1.error >= 1e-9 and abs(t2.error - t1.error) <= 1e-6 * t1.error)
            good = reached and abs(t2.iters - t1.iters) <= slack and err2 < 1e-6
            ok &= good
            print("[rank %d/%d] %-22s %-8s iters %d (single-domain %d%s) err %.2e rel-x-diff %.2e  %s"
                  % (rank, world, name, solver, t2.iters, t1.iters, "" if o_it is None else ", oracle %d" % o_it, t2.error, err2, "OK" if good else "FAIL"), flush=True)
        # ---- the slab stored as SELL-32 (set_format): product bit-identical to the single-domain SELL product, same CG ----
        D.set_format("sell", 32)
        S_full = full.to_sell(32)
        dyf = be.zeros(n)
        S_full.spmv(be.array(x), dyf)
        for _ in range(2):
            D.spmv(dx, dy)
        same_s = np.array_equal(dy.download(), dyf.download()[rb:re_])
        dxf = be.zeros(n)
        t1 = pkg.SolverTag(tol=1e-9, max_iterations=2000).solve("cg", S_full, dbf, dxf)
        dsol3 = be.zeros(re_ - rb)
        t2 = D.cg(db, dsol3, pkg.SolverTag(tol=1e-9, max_iterations=2000))
        xs, xf = dsol3.download(), dxf.download()[rb:re_]
        err3 = np.linalg.norm(xs - xf) / max(np.linalg.norm(xf), 1e-300)
        t3 = D.bicgstab(db, dsol3, pkg.SolverTag(tol=1e-9, max_iterations=2000))
        good = same_s and abs(t2.iters - t1.iters) <= 2 and err3 < 1e-6 and t3.error < 1e-9
        ok &= good
        print("[rank %d/%d] %-22s SELL-32 slab: spmv bit-exact=%s  cg iters %d (single-domain SELL %d) rel-x-diff %.2e  bicgstab iters %d  %s"
              % (rank, world, name, same_s, t2.iters, t1.iters, err3, t3.iters, "OK" if good else "FAIL"), flush=True)
        D.set_format("csr")
        # budget exhaustion must report the same iterate on every partitioning
        tag = D.cg(db, dsol, pkg.SolverTag(tol=1e-30, max_iterations=9))
        ref9 = o.cg(A, b, tol=1e-30, maxit=9)
        good = tag.iters == 9 and np.allclose(dsol.download(), ref9["x"][rb:re_], rtol=1e-9, atol=1e-12)
        ok &= good
        if not good:
            print("[rank %d] %s maxit case FAIL iters=%d" % (rank, name, tag.iters), flush=True)
        D.close()
    if "--big" in sys.argv:
        ok &= big_case(pkg, be, rank, world)
    if world > 1:
        t = torch.tensor([1.0 if ok else 0.0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item() > 0.5)
    be.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK", "PASS" if ok else "FAIL", "world", world, flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
