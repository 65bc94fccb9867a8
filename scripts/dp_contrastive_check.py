"""Real multi-GPU check of exchange 1: under torchrun, every rank owns N/W rows of a global batch;
InfoNCE / ConR / SupCon with `dp=DataParallelCtx()` (NCCL all-gather of the operands and of the row statistics) and
FDS with all-reduced epoch statistics must reproduce the single-process results on the concatenated batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/dp_contrastive_check.py
Prints one line per check on rank 0 and exits non-zero on a mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmdti_b200  # noqa: E402
from mmdti_b200.dist import DataParallelCtx  # noqa: E402
from mmdti_b200.models.contrastive import CT_Regress, CT_Single  # noqa: E402
from mmdti_b200.models.fds import FDS  # noqa: E402
from mmdti_b200.models.infonce import info_nce  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    dp = DataParallelCtx(grad_scale=1.0)
    N, D = 1024, 512
    M = N // world
    g = torch.Generator().manual_seed(0)
    f, f2 = torch.randn(N, D, generator=g), torch.randn(N, D, generator=g)
    y = torch.randn(N, 1, generator=g)
    yhat = y + 0.3 * torch.randn(N, 1, generator=g)
    w = torch.rand(N, generator=g) + 0.5
    cls = torch.randint(0, 10, (N, 1), generator=g)
    sl = slice(rank * M, rank * M + M)
    ok = True
    for mode, ltol, gtol in (("fp32", 1e-5, 1e-4), ("bf16", 1e-5, 1e-4)):        # same kernels on both sides: tight
        with mmdti_b200.precision(act=mode):
            a, b = f.cuda().requires_grad_(True), f2.cuda().requires_grad_(True)
            ref = [info_nce(a, b), CT_Regress(a, y.cuda(), yhat.cuda(), weights=w.cuda()), CT_Single(a, cls.cuda(), None)]
            sum(ref).backward()
            al, bl = f[sl].cuda().requires_grad_(True), f2[sl].cuda().requires_grad_(True)
            got = [info_nce(al, bl, dp=dp), CT_Regress(al, y[sl].cuda(), yhat[sl].cuda(), weights=w[sl].cuda(), dp=dp),
                   CT_Single(al, cls[sl].cuda(), None, dp=dp)]
            sum(got).backward()
        le = max(abs(x.item() - r.item()) / abs(r.item()) for x, r in zip(got, ref))
        ge = max(((al.grad - a.grad[sl]).norm() / a.grad[sl].norm()).item(), ((bl.grad - b.grad[sl]).norm() / b.grad[sl].norm()).item())
        errs = torch.tensor([le, ge], device="cuda")
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("contrastive %s world=%d: loss err %.2e, grad err %.2e" % (mode, world, errs[0].item(), errs[1].item()), flush=True)
        ok = ok and errs[0].item() < ltol and errs[1].item() < gtol
    # FDS epoch statistics
    nb = 20
    labels = torch.randn(N, generator=g)
    feats = torch.randn(N, 128, generator=g) * 2 + 0.5

    def mk():
        m = FDS(feature_dim=128, raw_data=np.array([0.0, 1.0]), col_data=None, using_scale=False, bucket_num=nb).cuda()
        m.min_value, m.bin_width = -2.0, 4.0 / nb
        return m

    ref = mk()
    ref.update_running_stats(feats.cuda(), labels.cuda(), 0)
    mine = mk()
    mine.dp = dp
    mine.update_running_stats(feats[sl].cuda(), labels[sl].cuda(), 0)
    e = max(((getattr(mine, k) - getattr(ref, k)).abs().max() / getattr(ref, k).abs().max()).item()
            for k in ("running_mean", "running_var", "num_samples_tracked"))
    et = torch.tensor([e], device="cuda")
    dist.all_reduce(et, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("fds world=%d: buffer err %.2e" % (world, et.item()), flush=True)
    ok = ok and et.item() < 1e-5
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print("DP CHECK", "OK" if ok else "FAILED", flush=True)
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
