"""ctypes binding of libmmdti_b200.so (the C ABI declared in include/mmdti_b200.h).

The product path has NO fallback: if the library is missing or a call fails, this raises."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmmdti_b200.so")

F32, BF16, F16 = 0, 1, 2
DTYPE_CODE = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}

_lib = None
launch_count = 0          # number of kernels-launching C-ABI calls made (bench.py reports it)


class MMDTIError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MMDTIError(
                "libmmdti_b200.so not found at %s — build it with `python mm-dti_b200/build.py` "
                "(there is no CPU / PyTorch fallback for the hot path)" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.mmdti_last_error.restype = ctypes.c_char_p
        _lib.mmdti_version.restype = ctypes.c_int
    return _lib


def _arg(a):
    if a is None:
        return ctypes.c_void_p(0)
    if torch.is_tensor(a):
        return ctypes.c_void_p(a.data_ptr())
    return a


_timeline = None          # when a list: (name, start_event, end_event) per call (bench.py's live kernel timing)


def start_timeline():
    global _timeline
    _timeline = []


def stop_timeline():
    """-> {name: (n_calls, total_seconds)}; call after torch.cuda.synchronize()."""
    global _timeline
    tl, _timeline = _timeline or [], None
    out = {}
    for name, a, b in tl:
        n, t = out.get(name, (0, 0.0))
        out[name] = (n + 1, t + a.elapsed_time(b) * 1e-3)
    return out


def call(name, *args):
    """Invoke an `int mmdti_*(...)` entry point; tensors -> device pointers; raises on error."""
    global launch_count
    fn = getattr(lib(), name)
    cargs = [_arg(a_) for a_ in args]
    if _timeline is not None:
        # events hug the launch: arguments are marshalled first so that host-side work is not inside the bracket
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
    rc = fn(*cargs)
    if rc != 0:
        raise MMDTIError("%s failed (%d): %s" % (name, rc, lib().mmdti_last_error().decode()))
    if _timeline is not None:
        b.record()
        _timeline.append((name, a, b))
    launch_count += 1


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def i32(x):
    return ctypes.c_int(int(x))


def i64(x):
    return ctypes.c_int64(int(x))


def u64(x):
    return ctypes.c_uint64(int(x) & 0xFFFFFFFFFFFFFFFF)


def f32(x):
    return ctypes.c_float(float(x))


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise MMDTIError("mmdti_b200 kernels need CUDA tensors (got a %s tensor); "
                             "there is no CPU fallback" % t.device)
