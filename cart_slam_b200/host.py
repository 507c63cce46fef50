"""ctypes binding of libcartb200_host.so - the C++ module layer (System / SystemModule / DataSource /
JSON config) that mirrors CART-SLAM's plugin surface on top of the cartb200 C ABI."""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcartb200_host.so")


class HostError(RuntimeError):
    pass


def _load():
    if not os.path.exists(_LIB_PATH):
        raise ImportError(f"{_LIB_PATH} is missing: build it with `python __graft_entry__.py`")
    C.CDLL(os.path.join(_HERE, "libcartb200.so"), mode=C.RTLD_GLOBAL)
    lib = C.CDLL(_LIB_PATH)
    lib.cartb200_host_last_error.restype = C.c_char_p
    lib.cartb200_host_run_config.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


_lib = _load()


def run_config(modules, left, right, skip_out_of_scope=False, sequential=True, want_labels=False, want_disparity=False):
    """Runs a reference-style module list (python list / JSON text) over host frames [n,H,W,3] uint8.
    Returns dict(planes=..., labels=..., disparity=...)."""
    text = modules if isinstance(modules, str) else json.dumps(modules)
    left = np.ascontiguousarray(left, dtype=np.uint8)
    right = np.ascontiguousarray(right, dtype=np.uint8)
    n, H, W, _ = left.shape
    planes = np.full((n, H, W), 255, np.uint8)
    labels = np.zeros((n, H, W), np.uint16) if want_labels else None
    disp = np.zeros((n, H, W), np.int16) if want_disparity else None
    rc = _lib.cartb200_host_run_config(text.encode(), int(skip_out_of_scope), W, H, n, left.ctypes.data, right.ctypes.data,
                                       int(sequential), planes.ctypes.data,
                                       labels.ctypes.data if labels is not None else None,
                                       disp.ctypes.data if disp is not None else None)
    if rc != 0:
        raise HostError(_lib.cartb200_host_last_error().decode())
    return dict(planes=planes, labels=labels, disparity=disp)
