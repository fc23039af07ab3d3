"""Per-kernel durations of the row-partitioned CG / SpMV on N GPUs (one process per GPU) via torch.profiler (CUPTI):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 profiles/dist_trace.py [n1] [iters]
Prints, for rank 0 and the last rank, the average duration of every kernel and the busy/idle split of the timeline.
Profiling aid only -- never a bench number."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import bench_workloads as bw  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from torch.profiler import profile, ProfilerActivity
    n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 64
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
    rows = n1 ** 3
    rb, re_ = bw.partition_rows(rows, world, rank)
    A = pkg.CsrMatrix.stencil(be, n1, n1, n1, row_begin=rb, row_end=re_)
    n = A.rows
    b, x = be.empty(n), be.zeros(n)
    be.check(be.L.ViennaCLCUDADfill_uniform(be.h, n, b.ptr, 0, 0, 1.0, 1.0))
    if world > 1:
        D = pkg.DistCsr(be, rows, rb, re_, A)
        solve = lambda t: D.cg(b, x, t)
        spmv = lambda: D.spmv(b, x)
        if rank == 0:
            print("dist info", D.info(), flush=True)
    else:
        solve = lambda t: t.solve("cg", A, b, x)
        spmv = lambda: A.spmv(b, x)
    solve(pkg.SolverTag(tol=1e-30, max_iterations=8))
    be.sync()
    if world > 1:
        dist.barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        solve(pkg.SolverTag(tol=1e-30, max_iterations=iters))
        be.sync()
        for _ in range(20):
            spmv()
        be.sync()
        if world == 1 and n1 <= 256:
            S = A.to_sell(32)
            pkg.SolverTag(tol=1e-30, max_iterations=iters).solve("cg", S, b, x)
            for _ in range(20):
                S.spmv(b, x)
            be.sync()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    if rank in (0, world - 1):
        by = {}
        for e in ev:
            by.setdefault(e.name[:70], []).append(e.time_range.end - e.time_range.start)
        t0, t1 = ev[0].time_range.start, ev[-1].time_range.end
        busy = sum(e.time_range.end - e.time_range.start for e in ev)
        out = ["[rank %d] timeline %.3f ms, sum of kernels %.3f ms" % (rank, (t1 - t0) / 1e3, busy / 1e3)]
        for k, v in sorted(by.items(), key=lambda kv: -sum(kv[1])):
            out.append("  %-70s n=%4d avg %9.2f us  min %9.2f  max %9.2f" % (k, len(v), np.mean(v), np.min(v), np.max(v)))
        print("\n".join(out), flush=True)
    if world > 1:
        dist.barrier()
        D.close()
    be.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
