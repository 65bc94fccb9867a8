"""Separates the mainloop rate from the epilogue cost of csrc/gemm_tc.cu: the plain-store GEMM at K = 64 (epilogue
only), the model's K, and a long K (asymptotic mainloop rate).  usage: python scripts/probe_gemm2.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmdti_b200  # noqa: E402,F401
from mmdti_b200 import ops_gemm  # noqa: E402

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
warm = len(sys.argv) > 1 and sys.argv[1] == "warm"


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if not warm:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


M = 8448
for N in (1536, 512, 2048):
    for K in (64, 512, 2048, 8192):
        x = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        b = torch.zeros(N, device=dev).bfloat16()
        y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        t_own = timeit(lambda: ops_gemm.gemm_bias(x, w, None, out=y))
        t_b = timeit(lambda: ops_gemm.gemm_bias(x, w, b, out=y))
        t_lib = timeit(lambda: torch.mm(x, w.t(), out=y))
        fl = 2.0 * M * N * K
        print("M=%d N=%4d K=%4d  own(store) %7.1f us %6.0f TF/s   own(bias) %7.1f us   lib %7.1f us %6.0f TF/s"
              % (M, N, K, t_own, fl / t_own / 1e6, t_b, t_lib, fl / t_lib / 1e6), flush=True)
