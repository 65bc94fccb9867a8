"""Raw C-ABI timing of K1/K2 (no autograd / allocation in the timed region): N back-to-back
launches over rotating buffer sets larger than L2, CUDA events around the whole batch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmdti_b200  # noqa: E402
from mmdti_b200 import _lib  # noqa: E402
from mmdti_b200._lib import DTYPE_CODE, call, f32, i32, i64, stream_ptr, u64  # noqa: E402

PEAK = 6551.0


def bench(launch, nsets, reps=5):
    for s in range(nsets):
        launch(s)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    n = 0
    for _ in range(reps):
        for s in range(nsets):
            launch(s)
            n += 1
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / n


def main():
    H, D = 64, 512
    only = os.environ.get("ONLY", "")
    for (B, L) in ((128, 66), (32, 258)):
        for pair in ("bf16", "fp32"):
            pdt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[pair]
            nsets = 4
            sets = []
            from mmdti_b200 import ops
            Lp = ops.pair_ld(L)
            for s in range(nsets):
                qkv = (torch.randn(B * L, 3 * D, device="cuda") * 0.5).bfloat16()
                bias = ops.PairPadFn.apply(torch.randn(B * H, L, L, device="cuda"), B, H, L, pdt)
                sets.append(dict(qkv=qkv, bias=bias, out=torch.empty_like(bias), o=torch.empty(B * L, D, device="cuda", dtype=torch.bfloat16),
                                 do=torch.randn(B * L, D, device="cuda").bfloat16(),
                                 dp=torch.nn.functional.pad(torch.randn(B, H, L, L, device="cuda") * 0.1, (0, Lp - L)).to(pdt).contiguous(),
                                 dpi=torch.empty_like(bias), dqkv=torch.empty_like(qkv)))
            nel = B * H * L * L
            esz = sets[0]["bias"].element_size()
            for p in (0.0, 0.1):
                def fwd(s):
                    d = sets[s]
                    q = d["qkv"]
                    call("mmdti_pair_attn_fwd", q[:, :D], q[:, D:2 * D], q[:, 2 * D:], i64(3 * D), d["bias"], d["out"], d["o"], i64(D),
                         i32(B), i32(H), i32(L), f32(8 ** -0.5), f32(p), u64(1), i32(1), i32(DTYPE_CODE[pdt]), stream_ptr())
                t = bench(fwd, nsets)
                by = 2 * nel * esz + 4 * B * L * D * 2
                print("K2 fwd  B=%d L=%d pair=%s p=%.1f: %7.1f us %6.0f GB/s (%.2f)" % (B, L, pair, p, t * 1e6, by / t / 1e9, by / t / 1e9 / PEAK), flush=True)

                def bwd(s, with_dp=True):
                    d = sets[s]
                    q, g = d["qkv"], d["dqkv"]
                    call("mmdti_pair_attn_bwd", q[:, :D], q[:, D:2 * D], q[:, 2 * D:], i64(3 * D), d["out"], d["o"], d["do"], i64(D),
                         d["dp"] if with_dp else None, d["dpi"], g[:, :D], g[:, D:2 * D], g[:, 2 * D:], i64(3 * D), i32(B), i32(H), i32(L),
                         f32(8 ** -0.5), f32(p), u64(1), i32(1), i32(DTYPE_CODE[pdt]), i32(DTYPE_CODE[pdt]), stream_ptr())
                t = bench(bwd, nsets)
                by = 3 * nel * esz + 9 * B * L * D * 2
                print("K2 bwd  B=%d L=%d pair=%s p=%.1f: %7.1f us %6.0f GB/s (%.2f)" % (B, L, pair, p, t * 1e6, by / t / 1e9, by / t / 1e9 / PEAK), flush=True)
        # K1 fwd
        from mmdti_b200.data import synthetic_molecules
        torch.manual_seed(0)
        tokens, dist, et, _ = synthetic_molecules(B, L - 2, seed=1)
        dist, et = dist.cuda(), et.cuda()
        ps = [torch.rand(128, device="cuda") * 3, torch.rand(128, device="cuda") * 3, torch.ones(961, device="cuda"), torch.zeros(961, device="cuda"),
              torch.randn(128, 128, device="cuda") * 0.05, torch.zeros(128, device="cuda"), torch.randn(64, 128, device="cuda") * 0.05, torch.zeros(64, device="cuda")]
        from mmdti_b200 import ops
        outs = [torch.empty(B, H, L, ops.pair_ld(L), device="cuda", dtype=torch.bfloat16) for _ in range(4)]

        def k1(s):
            call("mmdti_pair_bias_fwd", dist, et, *ps, None, outs[s], i32(B), i32(L), i32(128), i32(64), i32(961), i32(1), i32(0), stream_ptr())
        t = bench(k1, 4)
        by = B * H * L * L * 2 + B * L * L * 12
        print("K1 fwd  B=%d L=%d bf16: %7.1f us %6.0f GB/s (%.2f)  %.1f TFLOP/s" % (B, L, t * 1e6, by / t / 1e9, by / t / 1e9 / PEAK, B * L * L * 49152 / t / 1e12), flush=True)
        if only == "small":
            break


if __name__ == "__main__":
    main()
