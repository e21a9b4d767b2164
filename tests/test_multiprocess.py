"""CPU, world_size 2, gloo: the N>1 path of bench.py.  Replicas are independent sequences ("replicas
only"), so the only cross-rank logic is the max-over-ranks timing reduction and the metric gather."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total_ms = 100.0 + 50.0 * rank          # rank 1 is the slow one
        e2e_ms = 120.0 + 10.0 * rank
        t, e, gathered = bench.reduce_over_ranks(total_ms, e2e_ms, {"rank": rank, "fps": 10.0 / (total_ms / 1e3),
                                                                   "surfels": 1000 + rank}, world, torch.device("cpu"))
        if rank == 0:
            out.put((t, e, gathered, bench.aggregate_value(world, 10, t)))
    finally:
        dist.destroy_process_group()


def test_two_rank_reduction_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    t, e, gathered, value = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t == 150.0 and e == 130.0                       # MAX over ranks
    assert [g["rank"] for g in gathered] == [0, 1] and gathered[1]["surfels"] == 1001
    assert abs(value - 2 * 10 / 0.150) < 1e-9              # both replicas' frames within the slowest rank's time


def test_single_rank_is_a_no_op():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.reduce_over_ranks(5.0, 6.0, {"rank": 0}, 1, torch.device("cpu")) == (5.0, 6.0, None)
