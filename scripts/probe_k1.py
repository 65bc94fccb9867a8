"""K1 probe: times mmdti_pair_bias_fwd / mmdti_pair_bias_bwd alone at the config-2 shape (B=128, L=66)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmdti_b200  # noqa: E402,F401
from mmdti_b200 import ops  # noqa: E402
from mmdti_b200._lib import call, i32, stream_ptr  # noqa: E402
from mmdti_b200.data import synthetic_molecules  # noqa: E402

B, n_atoms, K, H, E = int(os.environ.get("B", 128)), 64, 128, 64, 961
L = n_atoms + 2
Lp = ops.pair_ld(L)
tokens, dist, et, _ = synthetic_molecules(B, n_atoms, seed=1)
g = torch.Generator(device="cuda").manual_seed(0)
dist, et = dist.cuda(), et.cuda()
means, stds = torch.rand(K, device="cuda", generator=g) * 3, torch.rand(K, device="cuda", generator=g) * 3
mul, bias = torch.ones(E, device="cuda"), torch.zeros(E, device="cuda")
w1, b1 = torch.randn(K, K, device="cuda", generator=g) * 0.05, torch.zeros(K, device="cuda")
w2, b2 = torch.randn(H, K, device="cuda", generator=g) * 0.05, torch.zeros(H, device="cuda")
out = torch.empty(B, H, L, Lp, device="cuda", dtype=torch.bfloat16)
d_out = (torch.randn(B, H, L, Lp, device="cuda", generator=g) * 0.01).bfloat16()
grads = [torch.zeros(n, device="cuda") for n in (K, K, E, E, K * K, K, H * K, H)]
sp = stream_ptr()


def fwd():
    call("mmdti_pair_bias_fwd", dist, et, means, stds, mul, bias, w1, b1, w2, b2, None, out, i32(B), i32(L), i32(K), i32(H), i32(E),
         i32(1), i32(0), sp)


def bwd():
    call("mmdti_pair_bias_bwd", d_out, dist, et, means, stds, mul, bias, w1, b1, w2, *grads, i32(B), i32(L), i32(K), i32(H), i32(E),
         i32(1), sp)


for name, fn in (("K1 fwd", fwd), ("K1 bwd", bwd)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        fn()
    e.record()
    torch.cuda.synchronize()
    print("%s B=%d L=%d: %.1f us" % (name, B, L, a.elapsed_time(e) * 1e3 / 5), flush=True)
