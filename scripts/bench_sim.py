"""Config 5 (BASELINE.json): contrastive-loss microbench — InfoNCE / SupCon / ConR similarity sweep,
N = 1K..64K embeddings x D, forward and forward+backward, against the tensor-pipe roofline.
usage: python scripts/bench_sim.py [--d 512] [--nmax 65536] [--json out.json]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmdti_b200  # noqa: E402
from mmdti_b200.models import contrastive as ctm  # noqa: E402
from mmdti_b200.models import infonce as infm  # noqa: E402


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--d", type=int, default=512)
    ap.add_argument("--nmin", type=int, default=1024)
    ap.add_argument("--nmax", type=int, default=65536)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
    except Exception:
        peak = 1590.0
    D = args.d
    rows = []
    N = args.nmin
    while N <= args.nmax:
        g = torch.Generator(device="cuda").manual_seed(N)
        f = torch.randn(N, D, device="cuda", generator=g)
        f2 = torch.randn(N, D, device="cuda", generator=g)
        y = torch.randn(N, 1, device="cuda", generator=g)
        yhat = y + 0.3 * torch.randn(N, 1, device="cuda", generator=g)
        cls = torch.randint(0, 10, (N, 1), device="cuda", generator=g)
        iters = max(2, min(50, int(2e12 / (N * N * D))))
        cases = {
            "infonce": (lambda a: infm.info_nce(a, f2), 2 * 2, 2 * 4),      # two directions
            "supcon": (lambda a: ctm.CT_Single(a, cls, None), 2, 4),
            "conr": (lambda a: ctm.CT_Regress(a, y, yhat), 2, 4),
        }
        for name, (fn, ffl, bfl) in cases.items():
            a = f.clone().requires_grad_(True)

            def fwd():
                with torch.no_grad():
                    fn(a)

            def fwdbwd():
                a.grad = None
                fn(a).backward()

            tf = timeit(fwd, iters)
            tb = timeit(fwdbwd, iters)
            flf = ffl * N * N * D
            flb = (ffl + bfl) * N * N * D           # + recompute (2) + dA (2) per direction
            row = {"loss": name, "N": N, "D": D, "fwd_ms": tf * 1e3, "fwd_tflops": flf / tf / 1e12, "fwd_frac": flf / tf / 1e12 / peak,
                   "fwdbwd_ms": tb * 1e3, "fwdbwd_tflops": flb / tb / 1e12, "fwdbwd_frac": flb / tb / 1e12 / peak}
            rows.append(row)
            print("%-8s N=%6d D=%d  fwd %8.3f ms %7.1f TF/s (%.2f)   fwd+bwd %8.3f ms %7.1f TF/s (%.2f of %.0f)"
                  % (name, N, D, row["fwd_ms"], row["fwd_tflops"], row["fwd_frac"], row["fwdbwd_ms"], row["fwdbwd_tflops"],
                     row["fwdbwd_frac"], peak), flush=True)
        N *= 2
    if args.json:
        json.dump({"peak_bf16_tflops": peak, "rows": rows}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
