"""ctypes loader of oracle/_build/libedge_oracle.so (built by oracle/Makefile).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libedge_oracle.so")


def load():
    if not os.path.exists(_LIB):
        subprocess.run(["make", "-s", "-C", _HERE], check=True)
    lib = ctypes.CDLL(_LIB)
    lib.eel_oracle_canny_rgb_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 5
    lib.eel_oracle_canny_rgb_batch.restype = ctypes.c_int
    return lib


def canny_rgb(imgs, low=100, high=200):
    lib = load()
    imgs = np.ascontiguousarray(imgs)
    n, h, w, _ = imgs.shape
    out = np.empty((n, h, w), dtype=np.uint8)
    rc = lib.eel_oracle_canny_rgb_batch(imgs.ctypes.data, out.ctypes.data, n, h, w, low, high)
    assert rc == 0
    return out
