"""K2 probe: times mmdti_pair_attn_fwd / _bwd alone (CUDA events, L2 flushed between launches) at the
BASELINE config-2 shape (or --L/--B) and prints achieved algorithmic GB/s against the measured HBM peak.
usage: python scripts/probe_k2.py [--L 66] [--B 128] [--p 0.1] [--pair bf16] [--iters 20] [--only fwd|bwd]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmdti_b200  # noqa: E402,F401
from mmdti_b200 import ops  # noqa: E402
from mmdti_b200._lib import DTYPE_CODE, call, f32, i32, i64, stream_ptr, u64  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=66)
ap.add_argument("--B", type=int, default=128)
ap.add_argument("--p", type=float, default=0.1)
ap.add_argument("--pair", default="bf16")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--only", default="")
ap.add_argument("--noflush", action="store_true")
a = ap.parse_args()
PEAK = 6551.0
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
B, H, L, D = a.B, 64, a.L, 512
Lp = ops.pair_ld(L)
pdt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[a.pair]
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B * L, 3 * D, device="cuda", generator=g) * 0.5).bfloat16()
pair = torch.randn(B, H, L, Lp, device="cuda", generator=g).to(pdt)
pair[..., L:] = float("-inf")
pout = torch.empty_like(pair)
o = torch.empty(B * L, D, device="cuda", dtype=torch.bfloat16)
d_o = (torch.randn(B * L, D, device="cuda", generator=g) * 0.1).bfloat16()
dpo = (torch.randn(B, H, L, Lp, device="cuda", generator=g) * 0.01).to(pdt)
dpo[..., L:] = 0
dpi = torch.empty_like(pair)
dqkv = torch.empty_like(qkv)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
code, pcode = DTYPE_CODE[torch.bfloat16], DTYPE_CODE[pdt]
scale = 8 ** -0.5


def fwd():
    call("mmdti_pair_attn_fwd", qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], i64(3 * D), pair, pout, o, i64(D), i32(B), i32(H), i32(L),
         f32(scale), f32(a.p), u64(7), i32(code), i32(pcode), stream_ptr())


def bwd():
    call("mmdti_pair_attn_bwd", qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], i64(3 * D), pout, o, d_o, i64(D), dpo, dpi, dqkv[:, :D],
         dqkv[:, D:2 * D], dqkv[:, 2 * D:], i64(3 * D), i32(B), i32(H), i32(L), f32(scale), f32(a.p), u64(7), i32(code), i32(pcode),
         i32(pcode), stream_ptr())


def timeit(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(a.iters):
        if not a.noflush:
            flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


esz = pair.element_size()
nel = B * H * L * Lp
act = B * L * D * 2
fwd()
if a.only != "bwd":
    t = timeit(fwd)
    by = 2 * nel * esz + 4 * act
    print("K2 fwd B=%d L=%d pair=%s p=%.2f: %7.1f us %6.0f GB/s (%.3f of %.0f)" % (B, L, a.pair, a.p, t * 1e6, by / t / 1e9, by / t / 1e9 / PEAK, PEAK))
if a.only != "fwd":
    t = timeit(bwd)
    by = 3 * nel * esz + 9 * act
    print("K2 bwd B=%d L=%d pair=%s p=%.2f: %7.1f us %6.0f GB/s (%.3f of %.0f)" % (B, L, a.pair, a.p, t * 1e6, by / t / 1e9, by / t / 1e9 / PEAK, PEAK))
