"""CPU, world_size 2 over gloo: the multi-rank plumbing of bench.py (rank-distinct synthetic envs, barrier, MAX-over-ranks
timing, whole-job aggregation).  The imagination path has no data-path collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench

    obs, act = bench.rank_inputs(4, rank)
    ms = bench.max_over_ranks([10.0 + rank, 20.0 - rank], torch.device("cpu"))
    agg = bench.whole_job_value(frames_per_rank=8, world=world, ms=ms[0])
    q.put((rank, float(obs.sum()), ms, agg))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduction_and_sharding():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, sum0, ms0, agg0), (r1, sum1, ms1, agg1) = out
    assert sum0 != sum1                      # ranks imagine different envs
    assert ms0 == ms1 == [11.0, 20.0]        # MAX over ranks, identical everywhere
    assert abs(agg0 - 2 * 8 / 11e-3) < 1e-6  # whole-job frames/s = all ranks' frames / max time
