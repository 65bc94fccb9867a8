"""Stand-alone times of the tcgen05 projection GEMMs (csrc/gemm_tc.cu) at the config-2 shapes next to the library
GEMM (+ the elementwise kernels each fused epilogue replaces).  L2 flushed between calls.
usage: python scripts/probe_gemm.py [--rows 8448] [--iters 10]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmdti_b200  # noqa: E402,F401
from mmdti_b200 import ops, ops_gemm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=8448)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--only", default="")
a = ap.parse_args()
M, D, D3, Fd = a.rows, 512, 1536, 2048
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


def rnd(*s, scale=1.0, dt=torch.bfloat16):
    return (torch.randn(*s, device=dev, generator=g) * scale).to(dt)


flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / a.iters * 1e3      # us


x, h = rnd(M, D), rnd(M, D)
w_in, b_in = rnd(D3, D, scale=0.05), rnd(D3, scale=0.1)
w_out, b_out = rnd(D, D, scale=0.05), rnd(D, scale=0.1)
w1, b1 = rnd(Fd, D, scale=0.05), rnd(Fd, scale=0.1)
w2, b2 = rnd(D, Fd, scale=0.05), rnd(D, scale=0.1)
u = rnd(M, Fd)
z = rnd(M, Fd)
res = rnd(M, D, dt=torch.float32)
ln_w, ln_b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
dqkv = rnd(M, D3, scale=0.1)
stats = torch.stack([res.mean(1), torch.rsqrt(res.var(1, unbiased=False) + 1e-5)]).contiguous()
dw, db, dbias = (torch.zeros(D, device=dev) for _ in range(3))
dbf = torch.zeros(Fd, device=dev)
cases = {
    "in_proj   fwd  (8448x1536x512)": (lambda: ops_gemm.gemm_bias(x, w_in, b_in), lambda: torch.addmm(b_in, x, w_in.t()), 2 * M * D3 * D),
    "fc1+gelu  fwd  (8448x2048x512)": (lambda: ops_gemm.gemm_bias_gelu(x, w1, b1), lambda: torch.nn.functional.gelu(torch.addmm(b1, x, w1.t())), 2 * M * Fd * D),
    "out+dropres+LN (8448x512x512)": (lambda: ops_gemm.gemm_dropres_ln(x, w_out, b_out, res, ln_w, ln_b, 0.1, 5), lambda: torch.addmm(b_out, x, w_out.t()), 2 * M * D * D),
    "fc2+dropres+LN (8448x512x2048)": (lambda: ops_gemm.gemm_dropres_ln(u, w2, b2, res, ln_w, ln_b, 0.1, 5), lambda: torch.addmm(b2, u, w2.t()), 2 * M * D * Fd),
    "dgrad out      (8448x512x512)": (lambda: ops_gemm.gemm_dgrad(x, w_out), lambda: torch.mm(x, w_out), 2 * M * D * D),
    "dgrad fc2+gelu'(8448x2048x512)": (lambda: ops_gemm.gemm_dgrad_gelu(x, w2, z, dbf), lambda: torch.mm(x, w2), 2 * M * D * Fd),
    "dgrad fc2 * stored gelu'      ": (lambda: ops_gemm.gemm_dgrad_gelu(x, w2, z, dbf, z_is_grad=True), lambda: torch.mm(x, w2), 2 * M * D * Fd),
    "fc1+gelu+gelu' fwd            ": (lambda: ops_gemm.gemm_bias_gelu(x, w1, b1, store_grad=True), lambda: torch.addmm(b1, x, w1.t()), 2 * M * Fd * D),
    "dgrad fc1+LN'  (8448x512x2048)": (lambda: ops_gemm.gemm_dgrad_lnbwd(u, w1, res, stats, ln_w, res, dw, db, dbias, 0.1, 5), lambda: torch.mm(u, w1), 2 * M * D * Fd),
    "dgrad in+LN'   (8448x512x1536)": (lambda: ops_gemm.gemm_dgrad_lnbwd(dqkv, w_in, res, stats, ln_w, res, dw, db, dbias, 0.1, 5), lambda: torch.mm(dqkv, w_in), 2 * M * D * D3),
    "wgrad fc2      (512x2048x8448)": (lambda: ops_gemm.gemm_wgrad(x, u), lambda: torch.mm(x.t(), u, out_dtype=torch.float32), 2 * M * D * Fd),
    "wgrad fc1      (2048x512x8448)": (lambda: ops_gemm.gemm_wgrad(u, x), lambda: torch.mm(u.t(), x, out_dtype=torch.float32), 2 * M * D * Fd),
    "wgrad out      (512x512x8448)": (lambda: ops_gemm.gemm_wgrad(x, h), lambda: torch.mm(x.t(), h, out_dtype=torch.float32), 2 * M * D * D),
    "wgrad in       (1536x512x8448)": (lambda: ops_gemm.gemm_wgrad(dqkv, h), lambda: torch.mm(dqkv.t(), h, out_dtype=torch.float32), 2 * M * D * D3),
}
print("%-34s %10s %10s %10s %10s" % ("case", "own us", "own TF/s", "lib us", "lib TF/s"))
for name, (own, lib, fl) in cases.items():
    if a.only and a.only not in name:
        continue
    t_own, t_lib = timeit(own), timeit(lib)
    print("%-34s %10.1f %10.0f %10.1f %10.0f   (library time = the bare GEMM, without the fused elementwise work)"
          % (name, t_own, fl / t_own / 1e6, t_lib, fl / t_lib / 1e6), flush=True)
