import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(os.path.abspath(__file__)) not in sys.path:
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def rel_err(a, b):
    """max |a-b| over finite entries / max |b|; also checks the non-finite pattern."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    fin = torch.isfinite(b)
    assert torch.equal(torch.isfinite(a), fin), "finiteness pattern differs"
    if (~fin).any():
        assert torch.equal(a[~fin], b[~fin]), "inf pattern differs"
    if not fin.any():
        return 0.0
    scale = max(b[fin].abs().max().item(), 1e-30)
    return (a[fin] - b[fin]).abs().max().item() / scale


def norm_err(a, b):
    """||a-b|| / ||b|| over finite entries."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    fin = torch.isfinite(b)
    return ((a[fin] - b[fin]).norm() / b[fin].norm().clamp_min(1e-30)).item()


@pytest.fixture
def report():
    """Appends diagnostic lines to gpurun_out/test_report.txt (comes back from the GPU box)."""
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)

    def _r(*parts):
        with open(os.path.join(d, "test_report.txt"), "a") as fh:
            fh.write(" ".join(str(p) for p in parts) + "\n")
    return _r
