"""Kernel-level timing probe (CUDA events, L2 flushed between launches) for K1/K2 at the
BASELINE config-2 shape.  Prints achieved GB/s against MEASURED_PEAKS.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmdti_b200  # noqa: E402
from mmdti_b200 import ops  # noqa: E402
from mmdti_b200.data import synthetic_molecules  # noqa: E402

PEAK = 6551.0
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, iters=10, warm=3):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    B, H = int(os.environ.get("B", 128)), 64
    for n_atoms in (64, 256):
        L = n_atoms + 2
        Bc = B if n_atoms == 64 else 32
        D = H * 8
        for pair in ("bf16", "fp16", "fp32"):
            pdt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[pair]
            qkv = (torch.randn(Bc * L, 3 * D, device="cuda") * 0.5).bfloat16()
            bias = torch.randn(Bc, H, L, L, device="cuda").to(pdt)
            o = torch.empty(Bc * L, D, device="cuda", dtype=torch.bfloat16)
            out = torch.empty_like(bias)
            esz = bias.element_size()
            nel = Bc * H * L * L
            for p in (0.0, 0.1):
                t = timeit(lambda: ops.PairAttnFn.apply(qkv, bias, Bc, H, L, 8 ** -0.5, p, 1, False))
                by = 2 * nel * esz + 4 * Bc * L * D * 2
                print("K2 fwd L=%d pair=%s p=%.1f: %.1f us  %.0f GB/s (%.2f of %.0f)" % (L, pair, p, t * 1e6, by / t / 1e9, by / t / 1e9 / PEAK, PEAK), flush=True)
            qkv_g = qkv.clone().requires_grad_(True)
            bias_g = bias.clone().requires_grad_(True)
            for p in (0.0, 0.1):
                oo, ss = ops.pair_attention(qkv_g, bias_g, Bc, H, L, 8 ** -0.5, p, 1)
                d_o, d_s = torch.randn_like(oo), torch.randn_like(ss)
                t = timeit(lambda: torch.autograd.grad([oo, ss], [qkv_g, bias_g], [d_o, d_s], retain_graph=True))
                by = 3 * nel * esz + 8 * Bc * L * D * 2
                print("K2 bwd L=%d pair=%s p=%.1f: %.1f us  %.0f GB/s (%.2f)" % (L, pair, p, t * 1e6, by / t / 1e9, by / t / 1e9 / PEAK), flush=True)
                t = timeit(lambda: torch.autograd.grad([oo], [qkv_g, bias_g], [d_o], retain_graph=True))
                by = 2 * nel * esz + 8 * Bc * L * D * 2
                print("K2 bwd(no dP') L=%d pair=%s p=%.1f: %.1f us  %.0f GB/s (%.2f)" % (L, pair, p, t * 1e6, by / t / 1e9, by / t / 1e9 / PEAK), flush=True)
            del qkv_g, bias_g, oo, ss
        # K1
        from mmdti_b200.models.encoder import GaussianLayer, NonLinearHead
        torch.manual_seed(0)
        gbf, proj = GaussianLayer(128, 961).cuda(), NonLinearHead(128, 64, "gelu").cuda()
        tokens, dist, et, _ = synthetic_molecules(Bc, n_atoms, seed=1)
        dist, et = dist.cuda(), et.cuda()
        for pair in ("bf16", "fp32"):
            with mmdti_b200.precision(act="bf16", pair=pair):
                t = timeit(lambda: proj(gbf(dist, et)))
                esz = 2 if pair == "bf16" else 4
                by = Bc * H * L * L * esz + Bc * L * L * 12
                print("K1 fwd L=%d pair=%s: %.1f us  %.0f GB/s (%.2f)  %.1f TFLOP/s" % (L, pair, t * 1e6, by / t / 1e9, by / t / 1e9 / PEAK, Bc * L * L * 49152 / t / 1e12), flush=True)
                outb = proj(gbf(dist, et))
                g = torch.randn_like(outb)
                params = list(gbf.parameters()) + list(proj.parameters())
                t = timeit(lambda: torch.autograd.grad([outb], params, [g], retain_graph=True), iters=5)
                print("K1 bwd (interim, library GEMMs) L=%d pair=%s: %.1f us" % (L, pair, t * 1e6), flush=True)


if __name__ == "__main__":
    main()
