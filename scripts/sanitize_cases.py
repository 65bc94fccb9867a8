"""Smallest invocation of every kernel family of the hot path, for compute-sanitizer (scripts/sanitize.sh).
Each case runs forward + backward once on tiny shapes; the results are checked for finiteness only (parity is the
job of tests/ -m gpu)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmdti_b200  # noqa: E402
from mmdti_b200 import ops_gemm  # noqa: E402
from mmdti_b200.data import synthetic_molecules  # noqa: E402
from mmdti_b200.models.contrastive import CT_Regress, CT_Single  # noqa: E402
from mmdti_b200.models.encoder import UnimolEncoder  # noqa: E402
from mmdti_b200.models.fds import FDS  # noqa: E402
from mmdti_b200.models.infonce import InfoNCE  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def finite(*ts):
    for t in ts:
        assert torch.isfinite(t.float()).all()


# encoder (K1 tcgen05 forward, K1 backward, K2, fused tcgen05 GEMMs with all epilogues incl. the 2-CTA LayerNorm ones,
# the cross-layer chain), training mode (dropout on), bf16 and the fp32 validation kernels
tokens, dist, et, _ = synthetic_molecules(2, 9, seed=1, ragged=True)
for act, pair in (("bf16", "bf16"), ("fp32", "fp32")):
    m = UnimolEncoder(encoder_layers=2).to(dev).train()
    with mmdti_b200.precision(act=act, pair=pair):
        rep = m(tokens.to(dev), dist.to(dev), et.to(dev))
        rep.float().pow(2).mean().backward()
    torch.cuda.synchronize()
    finite(rep, *[p.grad for p in m.parameters() if p.grad is not None])
    print("encoder", act, "ok", flush=True)

# stand-alone GEMMs on a shape with partial tiles (M = 200, D = 64)
x = torch.randn(200, 64, device=dev).bfloat16()
w = (torch.randn(192, 64, device=dev) * 0.05).bfloat16()
b = torch.zeros(192, device=dev).bfloat16()
y = ops_gemm.gemm_bias(x, w, b)
dx = ops_gemm.gemm_dgrad(y, w)
dw = ops_gemm.gemm_wgrad(y, x)
torch.cuda.synchronize()
finite(y, dx, dw)
print("gemm ok", flush=True)

# contrastive head: tcgen05 similarity kernels (phase 1, fused phase 2, two-step backward) and the fp32 kernels
g = torch.Generator().manual_seed(3)
f = torch.randn(300, 512, generator=g).to(dev).requires_grad_(True)
y1 = torch.randn(300, 1, generator=g).to(dev)
cls = torch.randint(0, 4, (300, 1), generator=g).to(dev)
inf = InfoNCE(512, 512).to(dev)
for act in ("bf16", "fp32"):
    with mmdti_b200.precision(act=act):
        loss = CT_Regress(f, y1, y1 * 0.9, w=0.2) + CT_Single(f, cls, None) + inf(f.view(30, 10, 512), f.view(30, 10, 512) * 0.5)
        loss.backward()
    torch.cuda.synchronize()
    finite(loss, f.grad)
    print("contrastive", act, "ok", flush=True)

# FDS
fds = FDS(feature_dim=512, raw_data=np.array([0.0, 1.0]), col_data=None, using_scale=False, bucket_num=8).to(dev)
fds.min_value, fds.bin_width = -2.0, 0.5
feats = torch.randn(300, 512, device=dev)
fds.update_running_stats(feats, y1, 0)
fds.update_last_epoch_stats(1)
out = fds.smooth(feats.clone().requires_grad_(True) * 1.0, y1, 1)
out.sum().backward()
torch.cuda.synchronize()
finite(out)
print("fds ok", flush=True)
print("SANITIZE CASES OK")
