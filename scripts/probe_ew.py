"""Elementwise-kernel probe at config-2 sizes (rows = 8448): LayerNorm bwd (+dropout), GELU bwd, column sums.
CUDA events, L2 flushed between launches.  usage: python scripts/probe_ew.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmdti_b200  # noqa: E402,F401
from mmdti_b200._lib import call, f32, i32, i64, stream_ptr, u64  # noqa: E402

rows, D, F_ = 8448, 512, 2048
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(rows, D, device="cuda", generator=g)
dy = torch.randn(rows, D, device="cuda", generator=g).bfloat16()
add = torch.randn(rows, D, device="cuda", generator=g)
w = torch.randn(D, device="cuda", generator=g)
b = torch.randn(D, device="cuda", generator=g)
y = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
st = torch.empty(2, rows, device="cuda")
dx = torch.empty(rows, D, device="cuda")
da = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
acc = torch.zeros(4 * D, device="cuda")
z = torch.randn(rows, F_, device="cuda", generator=g).bfloat16()
du = torch.randn(rows, F_, device="cuda", generator=g).bfloat16()
dz = torch.empty_like(z)
dbf = torch.zeros(F_, device="cuda")
dqkv = torch.randn(rows, 3 * D, device="cuda", generator=g).bfloat16()
dbq = torch.zeros(3 * D, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
sp = stream_ptr()
call("mmdti_layernorm_fwd", x, w, b, y, st[0], st[1], i32(rows), i32(D), f32(1e-5), i32(1), sp)
cases = {
    "layernorm_fwd (26 MB)": (lambda: call("mmdti_layernorm_fwd", x, w, b, y, st[0], st[1], i32(rows), i32(D), f32(1e-5), i32(1), sp), 26e6),
    "layernorm_bwd (61 MB)": (lambda: call("mmdti_layernorm_bwd", dy, x, w, st[0], st[1], add, dx, acc[:D], acc[D:2 * D], i32(rows), i32(D), i32(1), sp), 60.6e6),
    "layernorm_bwd_dropout (69 MB)": (lambda: call("mmdti_layernorm_bwd_dropout", dy, x, w, st[0], st[1], add, dx, acc[:D], acc[D:2 * D], da, acc[2 * D:3 * D],
                                                    i32(rows), i32(D), f32(0.1), u64(5), i32(1), sp), 69.2e6),
    "gelu_bwd (104 MB)": (lambda: call("mmdti_gelu_bwd", du, z, dz, dbf, i32(rows), i32(F_), i32(1), sp), 103.8e6),
    "gelu_fwd (69 MB)": (lambda: call("mmdti_gelu_fwd", z, dz, i64(rows * F_), i32(1), sp), 69.2e6),
    "colsum (26 MB)": (lambda: call("mmdti_colsum", dqkv, dbq, i32(rows), i32(3 * D), i32(1), sp), 26e6),
}
for name, (fn, by) in cases.items():
    for _ in range(3):
        fn()
    ts = []
    for _ in range(15):
        flush.zero_()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e) * 1e-3)
    ts.sort()
    t = ts[len(ts) // 2]
    print("%-32s %7.1f us  %6.0f GB/s" % (name, t * 1e6, by / t / 1e9), flush=True)
