"""CPU, world_size 2, gloo: the data-parallel plumbing of mm-dti_b200/dist.py (the collectives the
N>1 path issues: rank-major all-gather of the contrastive operands, sum / max all-reduces of the
FDS statistics, bucketed gradient all-reduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import mmdti_b200  # noqa: F401
        from mmdti_b200.dist import DataParallelCtx, allreduce_grads, shard_rows
        dp = DataParallelCtx()
        assert (dp.rank, dp.world, dp.grad_scale) == (rank, world, float(world))
        # rank-major all-gather
        local = torch.arange(6, dtype=torch.float32).view(3, 2) + 100 * rank
        full = dp.all_gather_rows(local)
        want = torch.cat([torch.arange(6, dtype=torch.float32).view(3, 2) + 100 * r for r in range(world)])
        assert torch.equal(full, want)
        lab = dp.all_gather_rows(torch.tensor([rank, rank + 10]))
        assert lab.tolist() == [0, 10, 1, 11]
        # reductions
        assert dp.all_reduce_sum(torch.tensor([1.0 + rank])).item() == 3.0
        flags = torch.tensor([rank, 1 - rank, 0], dtype=torch.int32)
        assert dp.all_reduce_max_(flags).tolist() == [1, 1, 0]
        assert shard_rows(8, rank, world) == (4 * rank, 4 * rank + 4)
        # bucketed gradient all-reduce: average of per-rank gradients, several buckets, a None grad
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 4), torch.nn.Linear(4, 2))
        x = torch.randn(5, 8, generator=torch.Generator().manual_seed(rank))
        net[1](net[0](x)).sum().backward()                       # net[2] gets no gradient on purpose
        g_local = [None if p.grad is None else p.grad.clone() for p in net.parameters()]
        n = allreduce_grads(list(net.parameters()), average=True, bucket_bytes=256)
        assert n >= 2
        gathered = [None] * world
        dist.all_gather_object(gathered, g_local)
        for i, p in enumerate(net.parameters()):
            parts = [g[i] if g[i] is not None else torch.zeros_like(p) for g in gathered]
            assert torch.allclose(p.grad, sum(parts) / world, atol=1e-6), i
        q.put((rank, "ok"))
    except Exception as e:                                       # noqa: BLE001
        q.put((rank, "fail: %r" % (e,)))
    finally:
        dist.destroy_process_group()


def test_data_parallel_plumbing_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
