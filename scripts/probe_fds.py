"""FDS kernels at a size where they are bandwidth-bound (N = 65 536 samples x 512 features): update_running_stats (two
passes over the features) and smooth (read + write), CUDA-event times and achieved algorithmic GB/s against the measured
HBM peak.  usage: python scripts/probe_fds.py [--n 65536]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmdti_b200  # noqa: E402,F401
from mmdti_b200.models.fds import FDS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=65536)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
N, D, NB = a.n, 512, 30
dev = "cuda"
fds = FDS(feature_dim=D, raw_data=np.array([0.0, 1.0]), col_data=None, using_scale=False, bucket_num=NB).to(dev)
fds.min_value, fds.bin_width = -3.0, 0.2
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(N, D, device=dev, generator=g)
y = torch.randn(N, 1, device=dev, generator=g)
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
    return tot / a.iters * 1e-3


fds.update_running_stats(x, y, 0)
fds.update_last_epoch_stats(1)
t_stats = timeit(lambda: fds.update_running_stats(x, y, 1))
xs = x.clone()
t_smooth = timeit(lambda: fds.smooth(xs, y, 1))
b_stats, b_smooth = 2 * N * D * 4, 2 * N * D * 4          # stats: two reads of the features; smooth: read + write
print("FDS update_running_stats N=%d: %.1f us, %.0f GB/s algorithmic (%.2f of %.0f)" % (N, t_stats * 1e6, b_stats / t_stats / 1e9, b_stats / t_stats / 1e9 / PEAK, PEAK))
print("FDS smooth               N=%d: %.1f us, %.0f GB/s algorithmic (%.2f of %.0f)" % (N, t_smooth * 1e6, b_smooth / t_smooth / 1e9, b_smooth / t_smooth / 1e9 / PEAK, PEAK))
