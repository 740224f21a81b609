"""world_size-2 gloo tests (CPU) of the data-parallel host logic: flat-bucket gradient averaging and the
VQ-EMA statistics reduction hook (crank_b200/net/_dp.py)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from crank_b200.net import _dp

    assert _dp.world_size() == 1 and _dp.stats_reducer() is None     # not enabled yet
    _dp.enable()
    assert _dp.active() and _dp.world_size() == world
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2, 3)), torch.nn.Parameter(torch.zeros(4))]
    params[0].grad = torch.full((5,), float(rank + 1))
    params[1].grad = torch.arange(6, dtype=torch.float32).view(2, 3) * (rank + 1)
    # params[2] has no gradient on any rank (like the EMA codebook): must be skipped consistently
    _dp.average_gradients(params)
    expect0 = torch.full((5,), (1 + 2) / 2.0)
    expect1 = torch.arange(6, dtype=torch.float32).view(2, 3) * 1.5
    ok = torch.allclose(params[0].grad, expect0) and torch.allclose(params[1].grad, expect1) and params[2].grad is None
    stats = torch.full((7,), float(rank + 1))
    _dp.stats_reducer()(stats)
    ok = ok and torch.allclose(stats, torch.full((7,), 3.0))
    _dp.disable()
    ok = ok and not _dp.active()
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_dp_gradient_bucket_and_stats_reduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)], res
