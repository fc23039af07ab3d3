"""world_size-2 gloo test (CPU) of the host-side multi-rank logic bench.py relies on: contiguous row partition,
broadcast of the 128-byte communicator id from rank 0, max-over-ranks timing reduction."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import bench
    import bench_workloads as bw
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert bench.dist_env() == (rank, world, rank)
    ids = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    assert ids[0] == bytes(range(128))
    rows = 512 ** 3
    rb, re_ = bw.partition_rows(rows, world, rank)
    t = torch.tensor([rb, re_], dtype=torch.int64)
    allr = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allr, t)
    assert allr[0][0].item() == 0 and allr[-1][1].item() == rows
    for a, b in zip(allr[:-1], allr[1:]):
        assert a[1].item() == b[0].item()
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    assert ms.item() == 10.0 + world - 1
    # uneven division still covers everything exactly once
    tot = torch.tensor([bw.partition_rows(1000003, world, rank)[1] - bw.partition_rows(1000003, world, rank)[0]], dtype=torch.int64)
    dist.all_reduce(tot)
    assert tot.item() == 1000003
    dist.barrier()
    dist.destroy_process_group()
    out.put(rank)


def test_two_rank_host_logic_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get() for _ in range(2)) == [0, 1]
