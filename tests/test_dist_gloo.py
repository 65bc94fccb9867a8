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
        # overlapped (hook-driven) reducer: same averaged gradients, buckets launched from gradient hooks
        from mmdti_b200.dist import OverlappedGradReducer
        torch.manual_seed(0)
        net2 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 4), torch.nn.Linear(4, 2))
        red = OverlappedGradReducer(net2.parameters(), average=True, bucket_bytes=256)
        assert len(red.buckets) >= 2
        for it in range(2):                                       # two steps: the reducer re-arms itself
            for prm in net2.parameters():
                prm.grad = None
            net2[1](net2[0](x * (it + 1))).sum().backward()       # net2[2] gets no gradient: finish() sends zeros
            g_loc = [None if prm.grad is None else None for prm in net2.parameters()]
            red.finish()
            ref_net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 4), torch.nn.Linear(4, 2))
            ref_net.load_state_dict(net2.state_dict())
            tot = [torch.zeros_like(prm) for prm in ref_net.parameters()]
            for r in range(world):
                xr = torch.randn(5, 8, generator=torch.Generator().manual_seed(r)) * (it + 1)
                for prm in ref_net.parameters():
                    prm.grad = None
                ref_net[1](ref_net[0](xr)).sum().backward()
                for t_, prm in zip(tot, ref_net.parameters()):
                    if prm.grad is not None:
                        t_ += prm.grad
            for t_, prm in zip(tot, net2.parameters()):
                assert torch.allclose(prm.grad, t_ / world, atol=1e-6)
        assert red.collectives == 2 * len(red.buckets)
        red.remove()
        # flat-bucket mode with a tail bucket: p.grad stays local, reduced_grad(p) * grad_scale is the average
        net3 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 4))
        net3.load_state_dict({k: v for k, v in net2.state_dict().items() if not k.startswith("2.")})
        red3 = OverlappedGradReducer(net3.parameters(), average=True, bucket_bytes=256, tail_params=[net3[0].bias], keep_flat=True)
        assert red3.buckets[-1] == [net3[0].bias] and all(net3[0].bias is not p_ for b in red3.buckets[:-1] for p_ in b)
        for prm in net3.parameters():
            prm.grad = None
        net3(x).sum().backward()
        local = [prm.grad.clone() for prm in net3.parameters()]
        red3.finish()
        tot = [torch.zeros_like(prm) for prm in net3.parameters()]
        for r in range(world):
            xr = torch.randn(5, 8, generator=torch.Generator().manual_seed(r))
            ref3 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 4))
            ref3.load_state_dict(net3.state_dict())
            ref3(xr).sum().backward()
            for t_, prm in zip(tot, ref3.parameters()):
                t_ += prm.grad
        for t_, prm, lg in zip(tot, net3.parameters(), local):
            assert torch.allclose(red3.reduced_grad(prm) * red3.grad_scale, t_ / world, atol=1e-6)
            assert torch.equal(prm.grad, lg)
            assert red3.reduced_grad(prm).data_ptr() % 16 == 0
        red3.remove()
        # bf16 wire format: the bucket travels as bf16 (half the bytes), the reduced gradient comes back as fp32 and equals
        # the sum of the bf16-rounded local gradients up to one bf16 rounding of the sum
        red4 = OverlappedGradReducer(net3.parameters(), average=True, bucket_bytes=256, keep_flat=True, comm_dtype=torch.bfloat16)
        for prm in net3.parameters():
            prm.grad = None
        net3(x).sum().backward()
        red4.finish()
        for t_, prm in zip(tot, net3.parameters()):
            got = red4.reduced_grad(prm) * red4.grad_scale
            assert got.dtype == torch.float32
            assert torch.allclose(got, t_ / world, rtol=2e-2, atol=2e-2 * float(t_.abs().max() / world) + 1e-6)
        red4.remove()
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
