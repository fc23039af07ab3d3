#!/usr/bin/env python
"""bench.py -- headline benchmark of the SpMV + Krylov hot path (contract: see the task statement / DESIGN.md section (d)).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload spmv|cg512]

Metric (BASELINE.json): "SpMV eff. GB/s (% HBM peak); CG iterations/sec at 1/2/4/8 B200".
  * default workload `spmv` = BASELINE configs[1]: y = A*x, 3-D 7-point Laplacian 256^3 per GPU (16.7M rows, 117M nnz, double
    CSR; SELL-32 reported beside it).  A step is one SpMV over the whole matrix.  With N > 1 GPUs the grid grows to
    256 x 256 x (256*N), row-partitioned one slab per rank with NVLink halo exchange (weak scaling).
    value = algorithmic bytes (12*nnz + 20*rows) of all ranks / time.
  * `cg512` = BASELINE configs[4]: pipelined CG on the 512^3 Laplacian, fixed budget of iterations, row-partitioned over
    N ranks (strong scaling); a step is one CG iteration; value = iterations / s.  The default run reports it under "cg".
Inputs are synthetic (device-side stencil generator; x = seeded uniform[1,2)) and far larger than the 126 MB L2, so no L2
flush is needed between steps.  Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max
over ranks.  One process per GPU (torchrun); rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from bench_workloads import WORKLOAD_SPMV, WORKLOAD_CG512, spmv_config      # noqa: E402  (shared by both arms)


# --------------------------------------------------------------------------------------------------------------------
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, device copy read+write)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.rows = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for nm, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------------------------------------------------
def host_cores():
    """Host threads this process may use.  NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its workers."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def ref_solve_c1(r, o, kind):
    """BASELINE configs[0] end to end on the host: solve(A, b, cg_tag(1e-8)) with host b in, host x out (examples/benchmarks/solver.cpp:106-120)."""
    A = o.stencil2d(1024, 1024)
    b = np.ones(A.rows)
    if kind == "reference":
        r.solve("cg", A, b, tol=1e-8, maxit=20)
        res = r.solve("cg", A, b, tol=1e-8, maxit=5000)
        sec = res["seconds"]
    else:
        t0 = time.perf_counter(); res = o.cg(A, b, tol=1e-8, maxit=5000); sec = time.perf_counter() - t0
    return {"workload": "solve(cg_tag(1e-8)) lap2d 1024^2, host b -> host x", "solve_ms": sec * 1e3, "iterations": int(res["iters"]),
            "iterations_per_sec": res["iters"] / sec, "error": float(res["error"])}


def run_reference(args):
    """--impl reference: the reference's own OpenMP host backend (oracle/_ref, compiled from /root/reference) timed on the
    host cores on the SAME workload / metric / config as the GPU arm, all host threads; rank 0 only."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    o = ol.oracle()
    kind = "reference" if ol.have_ref() else "port"
    r = ol.ref() if kind == "reference" else o
    cores = host_cores()
    o.set_threads(cores); r.set_threads(cores)
    n1 = 256
    ngpu = max(1, args.gpus)
    if args.workload == "spmv":
        # weak scaling: the GPU arm's matrix at N GPUs is the 256 x 256 x (256 N) grid -- the CPU arm runs the same matrix
        A = o.stencil3d(n1, n1, n1 * ngpu)
        x = o.uniform(A.cols, 1, 1.0, 2.0)
        steps, warm = max(1, args.steps), max(1, args.warmup)
        if kind == "reference":
            r.time_csr_spmv(A, x, warm)
            sec = r.time_csr_spmv(A, x, steps)
        else:
            y = np.zeros(A.rows)
            for _ in range(warm):
                o.csr_spmv(A, x, y)
            t0 = time.perf_counter()
            for _ in range(steps):
                o.csr_spmv(A, x, y)
            sec = time.perf_counter() - t0
        nbytes = 12 * A.nnz + 20 * A.rows
        val = nbytes * steps / sec / 1e9
        cfg = spmv_config(n1, ngpu)
        assert cfg["rows"] == A.rows and cfg["nnz"] == A.nnz
        cfg["note"] = "reference OpenMP host backend (viennacl/linalg/host_based), %d host threads, the whole %dx%dx%d matrix of the GPU arm" % (cores, n1, n1, n1 * ngpu)
        line = {"metric": "spmv_effective_GBps", "value": val, "unit": "GB/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
                "ms_per_step": sec / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "impl": "reference", "config": cfg,
                "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": kind,
                                 "sample": "%d SpMV passes over the full %dx%dx%d matrix" % (steps, n1, n1, n1 * ngpu)},
                "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        if ngpu == 1 and not args.no_extras:
            line["e2e_solve"] = ref_solve_c1(r, o, kind)
    else:
        # BASELINE configs[4] for real: pipelined CG on the 512^3 Laplacian (11.3 GB of host CSR), a bounded number of iterations
        A = o.stencil3d(512, 512, 512)
        b = np.ones(A.rows)
        its = max(2, min(args.steps, 20))
        if kind == "reference":
            res = r.solve("cg", A, b, tol=0.0, maxit=its)
            sec, done = res["seconds"], res["iters"]
        else:
            t0 = time.perf_counter(); res = o.cg(A, b, tol=0.0, maxit=its); sec = time.perf_counter() - t0; done = res["iters"]
        val = done / sec
        line = {"metric": "cg_iterations_per_sec", "value": val, "unit": "it/s", "n_gpus": args.gpus, "steps": done, "warmup": 0,
                "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "impl": "reference",
                "config": {"workload": WORKLOAD_CG512, "rows": A.rows, "nnz": A.nnz,
                           "note": "reference pipelined CG (cg.hpp:128-187, OpenMP host backend, %d threads) on the full 512^3 system, %d iterations incl. set-up" % (cores, done)},
                "cpu_baseline": {"value": val, "unit": "it/s", "cores": cores, "kind": kind, "sample": "%d CG iterations on the full 512^3 system" % done},
                "e2e": {"value": val, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------------
def cpu_baseline_spmv(n1, reps=10):
    """Reference OpenMP host backend on the box's host cores, bounded sample (rank 0, N = 1 only)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    o = ol.oracle()
    kind = "reference" if ol.have_ref() else "port"
    cores = host_cores()
    o.set_threads(cores)
    A = o.stencil3d(n1, n1, n1)
    x = o.uniform(A.cols, 1, 1.0, 2.0)
    if kind == "reference":
        r = ol.ref(); r.set_threads(cores)
        r.time_csr_spmv(A, x, 2)
        sec = r.time_csr_spmv(A, x, reps)
    else:
        y = np.zeros(A.rows)
        o.csr_spmv(A, x, y)
        t0 = time.perf_counter()
        for _ in range(reps):
            o.csr_spmv(A, x, y)
        sec = time.perf_counter() - t0
    nbytes = 12 * A.nnz + 20 * A.rows
    return {"value": nbytes * reps / sec / 1e9, "unit": "GB/s", "cores": cores, "kind": kind,
            "sample": "%d CSR SpMV passes over the full 256^3 matrix (reference host_based::prod_impl, OpenMP)" % reps,
            "ms_per_step": sec / reps * 1e3}


def cpu_baseline_cg(n1=1024, iters=200):
    """BASELINE configs[0] on the host: the reference's pipelined CG (cg.hpp:128-187, OpenMP host backend) on the 2-D 5-point
    Laplacian n1 x n1, fixed iteration budget, all host cores -- the CPU arm of the 'CG iterations/sec' half of the metric."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    o = ol.oracle()
    cores = host_cores()
    o.set_threads(cores)
    A = o.stencil2d(n1, n1)
    b = np.ones(A.rows)
    if ol.have_ref():
        r = ol.ref(); r.set_threads(cores)
        r.solve("cg", A, b, tol=1e-300, maxit=10)
        res = r.solve("cg", A, b, tol=1e-300, maxit=iters)
        sec, its, kind = res["seconds"], res["iters"], "reference"
    else:
        o.cg(A, b, tol=1e-300, maxit=10)
        t0 = time.perf_counter()
        res = o.cg(A, b, tol=1e-300, maxit=iters)
        sec, its, kind = time.perf_counter() - t0, res["iters"], "port"
    return {"iterations_per_sec": its / sec, "iterations": its, "cores": cores, "kind": kind,
            "sample": "%d pipelined CG iterations on the %dx%d Laplacian (reference cg.hpp, OpenMP host backend)" % (its, n1, n1)}


def legacy_cuda_baseline(n1=256):
    """Runs _legacy_cuda_child in a SEPARATE process (python bench.py --legacy-cuda-child): the legacy kernels live in their own
    CUDA context and a crash inside them (they are unmodified 2016 code on a 2025 GPU) cannot take the bench line down."""
    try:
        p = subprocess.run([sys.executable, "-X", "faulthandler", os.path.abspath(__file__), "--legacy-cuda-child", str(n1)],
                           capture_output=True, text=True, timeout=600)
    except subprocess.TimeoutExpired:
        return {"unavailable": "legacy CUDA baseline timed out after 600 s"}
    for ln in reversed(p.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
    return {"unavailable": "legacy CUDA baseline process died (rc %d): %s" % (p.returncode, p.stderr.strip()[-300:])}


def _legacy_cuda_child(n1=256):
    """SECOND baseline (rank 0, N = 1): the reference's OWN CUDA backend (viennacl/linalg/cuda/*, unmodified, compiled for sm_100
    by oracle/Makefile -> oracle/_ref/libvcl_ref_cuda.so) on the same GPU and the same matrices.  Every result is first checked
    against the reference host backend / the oracle (check K1's output before trusting it as a baseline)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    rc = ol.ref_cuda()
    if rc is None:
        return {"unavailable": "oracle/_ref/libvcl_ref_cuda.so not built or no GPU visible to it"}
    o = ol.oracle()
    o.set_threads(host_cores())
    out = {"what": "unmodified viennacl CUDA backend (cuda/sparse_matrix_operations.hpp, cuda/iterative_operations.hpp), nvcc -arch=sm_100, same GPU"}
    A = o.stencil3d(n1, n1, n1)
    x = o.uniform(A.cols, 1, 1.0, 2.0)
    y_ref = o.csr_spmv(A, x)
    nbytes = 12 * A.nnz + 20 * A.rows
    for fmt in ("csr", "sell"):
        try:
            reps = 20
            y, sec = rc.spmv(A, x, reps=reps, fmt=fmt)
            err = float(ol.rel_err(y, y_ref).max())
            out[fmt + "_spmv_256"] = {"ms_per_step": sec / reps * 1e3, "effective_GBps": nbytes * reps / sec / 1e9, "max_rel_err_vs_oracle": err,
                                      "correct": bool(err <= 1e-12)}
        except Exception as e:                                  # a crash of the legacy kernels must not take the bench line down
            out[fmt + "_spmv_256"] = {"failed": str(e)[:200]}
    del A, x, y_ref
    try:
        A = o.stencil2d(1024, 1024)
        b = np.ones(A.rows)
        rc.solve("cg", A, b, tol=1e-8, maxit=50)
        res = min((rc.solve("cg", A, b, tol=1e-8, maxit=5000) for _ in range(3)), key=lambda r_: r_["seconds"])    # best of 3: its per-iteration
        # blocking read-back makes single runs jitter by 4x on a busy host
        true = float(np.linalg.norm(b - o.csr_spmv(A, res["x"])) / np.linalg.norm(b))
        out["cg_lap2d_1024"] = {"iterations": res["iters"], "solve_ms": res["seconds"] * 1e3, "iterations_per_sec": res["iters"] / res["seconds"],
                                "error": res["error"], "true_residual": true, "correct": bool(true < 1e-7)}
    except Exception as e:
        out["cg_lap2d_1024"] = {"failed": str(e)[:200]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="spmv", choices=["spmv", "cg512"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the SELL / CG side measurements")
    ap.add_argument("--legacy-cuda-child", type=int, default=0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.legacy_cuda_child:
        print(json.dumps(_legacy_cuda_child(args.legacy_cuda_child)), flush=True)
        return
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.load_package()

    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

    be = pkg.Backend(local)             # raises without a B200: there is no CPU fallback
    if world > 1:
        ids = [be.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        be.comm_init(ids[0], rank, world)

    def barrier():
        be.sync()
        if world > 1:
            dist.barrier()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak, peak_src = load_peaks()
    n1 = 256
    line = None
    sampler = ClockSampler(local)

    if args.workload == "spmv":
        import bench_workloads as bw
        line = bw.spmv_workload(pkg, be, args, rank, world, n1, barrier, max_over_ranks, sampler, peak, peak_src)
        if rank == 0 and not args.no_extras:
            line["sell"] = bw.sell_side(pkg, be, args, n1, peak) if world == 1 else None
        if not args.no_extras:
            cg = bw.cg_side(pkg, be, args, rank, world, barrier, max_over_ranks)
            if rank == 0:
                line["cg"] = cg
        # clocks / throttle reasons were sampled (every 0.1 s) from the first warm-up step to here: the timed SpMV region, the e2e
        # region and the side measurements -- all of it GPU load of this process
        sampler.stop_flag.set()
        if rank == 0:
            line["clocks"] = sampler.summary()
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_spmv(n1)
            if not args.no_extras and line.get("cg") and "lap2d_1024" in line["cg"]:
                line["cg"]["lap2d_1024"]["cpu_baseline"] = cpu_baseline_cg()
            if not args.no_extras:
                line["legacy_cuda_baseline"] = legacy_cuda_baseline(n1)
                lg = line["legacy_cuda_baseline"]
                if lg.get("csr_spmv_256", {}).get("effective_GBps"):
                    lg["ratio_csr_spmv"] = line["value"] / lg["csr_spmv_256"]["effective_GBps"]
                if lg.get("sell_spmv_256", {}).get("effective_GBps") and line.get("sell"):
                    lg["ratio_sell_spmv"] = line["sell"]["value"] / lg["sell_spmv_256"]["effective_GBps"]
                if lg.get("cg_lap2d_1024", {}).get("iterations_per_sec") and line.get("cg"):
                    lg["ratio_cg_lap2d_1024"] = line["cg"]["lap2d_1024"]["iterations_per_sec"] / lg["cg_lap2d_1024"]["iterations_per_sec"]
                if lg.get("csr_spmv_256", {}).get("correct") is False:
                    lg["note"] = ("the legacy CSR kernel K1 (cuda/sparse_matrix_operations.hpp:137-178) returns WRONG rows on sm_100: its warp-synchronous "
                                  "shared-memory reduction has no __syncwarp (:168-173, SURVEY 0); its time is still reported, the SELL kernel and the CG "
                                  "(adaptive CSR kernel K7) are correct")
    else:
        import bench_workloads as bw
        line = bw.cg512_workload(pkg, be, args, rank, world, barrier, max_over_ranks, sampler, peak, peak_src)

    if rank == 0:
        print(json.dumps(line), flush=True)
    barrier()
    be.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
