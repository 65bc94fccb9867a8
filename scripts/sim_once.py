"""One forward + backward of a contrastive loss on the tensor-core path (for ncu captures)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmdti_b200
from mmdti_b200.models import infonce as infm, contrastive as ctm

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
which = sys.argv[3] if len(sys.argv) > 3 else "infonce"
g = torch.Generator(device="cuda").manual_seed(1)
a = torch.randn(N, D, device="cuda", generator=g, requires_grad=True)
b = torch.randn(N, D, device="cuda", generator=g)
y = torch.randn(N, 1, device="cuda", generator=g)
yh = y + 0.3 * torch.randn(N, 1, device="cuda", generator=g)
with mmdti_b200.precision(act="bf16"):
    for _ in range(2):
        a.grad = None
        loss = infm.info_nce(a, b) if which == "infonce" else ctm.CT_Regress(a, y, yh)
        loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
