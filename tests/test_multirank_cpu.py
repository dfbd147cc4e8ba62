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


def _ddp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from torch.nn.parallel import DistributedDataParallel as DDP

    from diamond_b200.utils import allreduce_gradients, broadcast_if_needed

    def make():
        torch.manual_seed(0)
        return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.SiLU(), torch.nn.Flatten(), torch.nn.Linear(8 * 36, 5))

    g = torch.Generator().manual_seed(100 + rank)            # every rank trains on its own shard
    x, y = torch.randn(4, 3, 6, 6, generator=g), torch.randn(4, 5, generator=g)
    ref = DDP(make())                                       # what the reference does (utils.py:105-106)
    torch.nn.functional.mse_loss(ref(x), y).backward()
    mine = make()
    torch.nn.functional.mse_loss(mine(x), y).backward()
    local = [p.grad.clone() for p in mine.parameters()]
    calls = allreduce_gradients(list(mine.parameters()), bucket_bytes=1024)   # small buckets: several collectives
    same = all(torch.allclose(a.grad, b.grad, rtol=1e-6, atol=1e-7) for a, b in zip(mine.parameters(), ref.module.parameters()))
    changed = any(not torch.equal(a, p.grad) for a, p in zip(local, mine.parameters()))
    (seed,) = broadcast_if_needed(1234 + rank)
    q.put((rank, same, changed, calls, seed))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_torch_ddp():
    """explicit bucketed gradient averaging == the DDP wrapper the reference uses."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, changed, calls, seed in out:
        assert same and changed
        assert calls >= 2                                    # 1 KB buckets: conv / linear weights go separately
        assert seed == 1234                                  # rank 0's object everywhere (utils.py:97-102)


def test_allreduce_gradients_is_a_noop_without_a_process_group():
    from diamond_b200.utils import allreduce_gradients

    lin = torch.nn.Linear(3, 2)
    lin(torch.ones(1, 3)).sum().backward()
    before = lin.weight.grad.clone()
    assert allreduce_gradients(list(lin.parameters())) == 0
    assert torch.equal(before, lin.weight.grad)


def _flat_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diamond_b200.utils import allreduce_native_gradients

    class Model(torch.nn.Module):   # what a native backward leaves behind: every .grad is a view of ONE flat buffer
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Parameter(torch.zeros(3, 4))
            self.b = torch.nn.Parameter(torch.zeros(5))
            self.frozen = torch.nn.Parameter(torch.zeros(2), requires_grad=False)

    m = Model()
    flat = torch.arange(17, dtype=torch.float32) * (rank + 1)        # rank-dependent gradients
    m.a.grad, m.b.grad = flat[0:12].view(3, 4), flat[12:17]
    m.last_flat_grad = flat
    calls_aliased = allreduce_native_gradients(m)
    want = torch.arange(17, dtype=torch.float32) * (sum(r + 1 for r in range(world)) / world)
    ok_aliased = torch.equal(flat, want) and torch.equal(m.a.grad.reshape(-1), want[:12]) and m.a.grad.data_ptr() == flat.data_ptr()
    # gradients that no longer alias the buffer (e.g. after gradient accumulation): bucketed fallback, same average
    m.a.grad = (torch.ones(3, 4) * (rank + 1)).clone()
    calls_fallback = allreduce_native_gradients(m)
    ok_fallback = torch.allclose(m.a.grad, torch.full((3, 4), sum(r + 1 for r in range(world)) / world)) and torch.allclose(m.b.grad, want[12:] * 1.0)
    q.put((rank, calls_aliased, ok_aliased, calls_fallback, ok_fallback))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_flat_buffer_allreduce_is_one_collective():
    """the training blocks of bench.py average a model's gradients with ONE all_reduce on the flat buffer the
    native backward filled (`allreduce_native_gradients`); when `.grad`s do not alias that buffer it falls back to buckets."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_flat_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, calls_aliased, ok_aliased, calls_fallback, ok_fallback in out:
        assert calls_aliased == 1 and ok_aliased
        assert calls_fallback >= 1 and ok_fallback
