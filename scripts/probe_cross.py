"""Stand-alone timing of the cross-modal fusion block at a bench shape: CrossAttentionModel forward +
backward + masked pooling, bf16 mode, dropout as configured by the reference (0.3 hidden / 0.2 attention).  CUDA events around
whole iterations with the L2 flushed in between; used under ncu for the per-kernel captures in profiles/.
    python scripts/probe_cross.py [--B 128] [--L1 66] [--L2 64] [--iters 10]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmdti_b200  # noqa: E402
from mmdti_b200.models.cross_modal import CrossAttentionModel, crossmodal_config, fuse_and_pool  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=128)
ap.add_argument("--L1", type=int, default=66)
ap.add_argument("--L2", type=int, default=64)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda")
torch.manual_seed(0)
net = CrossAttentionModel(crossmodal_config(), 1).to(dev).train()
x1 = torch.randn(a.B, a.L1, 512, device=dev, requires_grad=True)
x2 = torch.randn(a.B, a.L2, 512, device=dev, requires_grad=True)
m1 = torch.ones(a.B, a.L1, dtype=torch.bool, device=dev)
m2 = torch.arange(a.L2, device=dev)[None, :] < torch.randint(a.L2 // 2, a.L2 + 1, (a.B, 1), device=dev)
up = torch.randn(a.B, 512, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def it():
    t2g, g2t = net(x1, x2, m1, m2)
    (fuse_and_pool(t2g, g2t, m1, m2) * up).sum().backward()


with mmdti_b200.precision(act="bf16", pair="bf16"):
    for _ in range(3):
        it()
    ts = []
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        it()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
print("cross-modal block B=%d L1=%d L2=%d: fwd+bwd %.3f ms (median of %d, host-launched)" % (a.B, a.L1, a.L2, sorted(ts)[len(ts) // 2], a.iters))
