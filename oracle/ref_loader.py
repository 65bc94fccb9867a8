"""Import the reference's OWN hot-path files from /root/reference (read-only) so that
``make_golden.py`` can execute them.  TEST INFRASTRUCTURE; only usable in the build
container (the GPU box has no /root/reference) -- nothing under tests/ -m gpu, smoke()
or bench.py may call this.

How: the reference's ``models/__init__.py`` eagerly imports the whole trainer stack
(rdkit, ...), so a bare namespace module ``models`` whose ``__path__`` points at
/root/reference/models is registered instead; ``unicore`` and ``addict`` (both absent
in this image) resolve to ``oracle/shims``.  The reference's logger creates ``./logs``
in the cwd at import (utils/base_logger.py:44-48), so imports run from a temp cwd.
"""
import contextlib
import importlib
import os
import sys
import tempfile
import types

_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
# where the reference tree is looked for: an explicit override, the build container's read-only checkout, or the
# git-ignored copy that scripts/install_reference.sh makes under baseline/_ref so that bench.py --impl reference can
# run the reference's own modules on the GPU box (kind "reference").  The -m gpu tests and smoke() never come here.
_CANDIDATES = [os.environ.get("MMDTI_REFERENCE_ROOT"), "/root/reference",
               os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")]
REF_ROOT = next((c for c in _CANDIDATES if c and os.path.isdir(os.path.join(c, "models"))), "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "models"))


@contextlib.contextmanager
def _tmp_cwd():
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            yield
        finally:
            os.chdir(old)


_loaded = {}


def load():
    """Returns a dict of the reference's modules: transformers, infonce, contrastive,
    loss, fds, mm_model, util (utils.util)."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    for p in (_SHIMS, REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    launch_blocking = os.environ.get("CUDA_LAUNCH_BLOCKING")
    with _tmp_cwd():
        pkg = types.ModuleType("models")
        pkg.__path__ = [os.path.join(REF_ROOT, "models")]
        sys.modules["models"] = pkg
        names = ["transformers", "infonce", "contrastive", "loss", "fds", "mm_model"]
        for n in names:
            _loaded[n] = importlib.import_module("models." + n)
        _loaded["util"] = importlib.import_module("utils.util")
        # the reference logger writes ./logs/mm_dti_<ts>.log in the (temporary) cwd
        import logging
        lg = logging.getLogger("MM-DTI")
        for h in list(lg.handlers):
            if isinstance(h, logging.FileHandler):
                lg.removeHandler(h)
    # models/mm_model.py:7 force-sets CUDA_LAUNCH_BLOCKING=1 at import; undo it.
    if launch_blocking is None:
        os.environ.pop("CUDA_LAUNCH_BLOCKING", None)
    else:
        os.environ["CUDA_LAUNCH_BLOCKING"] = launch_blocking
    return _loaded
