"""ctypes binding of libeel.so -- prototypes are parsed from include/eel.h, so header and binding
cannot drift.  There is no CPU fallback: if the library is missing the import fails loudly.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, "..", "include", "eel.h")
LIB_PATH = os.path.join(_HERE, "libeel.so")

EEL_F32, EEL_BF16 = 0, 1

_CTYPES = {
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "size_t": ctypes.c_size_t,
    "float": ctypes.c_float,
    "eel_stream": ctypes.c_void_p,
}


def parse_header(path=HEADER):
    """[(name, restype, [argtypes])] for every prototype in eel.h."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    protos = []
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(eel_\w+)\s*\(([^;{}]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef"):
            continue
        if "*" in ret:
            restype = ctypes.c_char_p if "char" in ret else ctypes.c_void_p
        else:
            restype = _CTYPES[ret]
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                elif "unsigned long long" in a:
                    argtypes.append(ctypes.c_ulonglong)
                else:
                    ty = re.sub(r"\s+\w+$", "", a).replace("const ", "").strip()
                    argtypes.append(_CTYPES[ty])
        protos.append((name, restype, argtypes))
    return protos


class EelError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "eel_unet_b200: %s is missing -- build it with `python -m eel_unet_b200.build` "
            "(there is no CPU or PyTorch fallback for the kernels)" % LIB_PATH
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, restype, argtypes in parse_header():
        fn = getattr(lib, name)  # AttributeError if the header declares something the .so lacks
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = _load()


def check(rc, what=""):
    if rc != 0:
        raise EelError("%s failed (%d): %s" % (what or "eel call", rc, lib.eel_last_error().decode()))


_profile = None


def set_profiler(records):
    """records: a list that receives (name, args, start_event, end_event) per C-ABI call, or None to stop.
    Used by bench.py to time each kernel family with CUDA events on the launching stream."""
    global _profile
    _profile = records


def call(name, *args):
    if _profile is None:
        check(getattr(lib, name)(*args), name)
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    check(getattr(lib, name)(*args), name)
    e.record()
    _profile.append((name, args, s, e))


def dtype_code(t):
    if t.dtype == torch.float32:
        return EEL_F32
    if t.dtype == torch.bfloat16:
        return EEL_BF16
    raise EelError("unsupported storage dtype %s" % t.dtype)


def on_device(fn):
    """decorator of the public entry points: run ``fn`` with the device of its first CUDA tensor argument current, so that
    ``stream()`` / ``workspace()`` and the kernels' per-device attributes belong to the device the data lives on"""
    import functools

    @functools.wraps(fn)
    def wrapped(*a, **k):
        for t in list(a) + list(k.values()):
            if isinstance(t, torch.Tensor) and t.is_cuda:
                if t.device.index != torch.cuda.current_device():
                    with torch.cuda.device(t.device):
                        return fn(*a, **k)
                break
        return fn(*a, **k)

    return wrapped


def ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise EelError("eel_unet_b200 kernels need CUDA tensors (no CPU fallback); got a %s tensor" % t.device)
    if t.device.index != torch.cuda.current_device():
        # a launch on the current device's stream with another device's pointer would fault (or worse): refuse loudly.
        # The public entry points (model forward, loss, edges, data, metrics) switch devices themselves (on_device).
        raise EelError("tensor on %s but the current CUDA device is %d: wrap the call in torch.cuda.device(%r)"
                       % (t.device, torch.cuda.current_device(), str(t.device)))
    if not t.is_contiguous():
        raise EelError("non-contiguous tensor passed to a kernel")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


_ws = {}


def workspace(nbytes, device, slot=0):
    """Grow-only scratch buffer owned by the torch caching allocator (the library never allocates)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), slot,
           torch.cuda.current_stream().cuda_stream)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


def launch_count():
    return int(lib.eel_launch_count())
