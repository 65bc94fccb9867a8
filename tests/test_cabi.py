"""CPU: the C-ABI library loads and exports every symbol declared in include/mmdti_b200.h."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mmdti_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmdti_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import mmdti_b200  # noqa: F401
    from mmdti_b200 import _lib
    from mmdti_b200.build import build
    build(verbose=False)
    lib = _lib.lib()
    names = _declared()
    assert "mmdti_pair_attn_fwd" in names and len(names) >= 10
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.mmdti_version() >= 100
    assert isinstance(lib.mmdti_last_error(), bytes)


def test_product_path_fails_loudly_without_cuda():
    import pytest
    import torch
    from mmdti_b200 import ops
    from mmdti_b200._lib import MMDTIError
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(MMDTIError):
        ops.pair_attention(torch.zeros(10, 96), torch.zeros(1, 4, 10, 10), 1, 4, 10, 1.0)
