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
    lib.cartb200_host_decode_png.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.cartb200_host_inflate.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    lib.cartb200_host_run_source.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cartb200_host_run_config_ex.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


_lib = _load()


def run_config(modules, left, right, skip_out_of_scope=False, sequential=True, want_labels=False, want_disparity=False,
               Q=None, want_depth=False):
    """Runs a reference-style module list (python list / JSON text) over host frames [n,H,W,3] uint8.
    Q: optional 4x4 reprojection matrix of the source (CameraIntrinsics::Q) for the "depth" module.
    Returns dict(planes=..., labels=..., disparity=..., depth=...)."""
    text = modules if isinstance(modules, str) else json.dumps(modules)
    left = np.ascontiguousarray(left, dtype=np.uint8)
    right = np.ascontiguousarray(right, dtype=np.uint8)
    n, H, W, _ = left.shape
    planes = np.full((n, H, W), 255, np.uint8)
    labels = np.zeros((n, H, W), np.uint16) if want_labels else None
    disp = np.zeros((n, H, W), np.int16) if want_disparity else None
    depth = np.zeros((n, H, W, 3), np.float32) if want_depth else None
    q = np.ascontiguousarray(np.asarray(Q, np.float32).reshape(16)) if Q is not None else None
    rc = _lib.cartb200_host_run_config_ex(text.encode(), int(skip_out_of_scope), W, H, n, left.ctypes.data, right.ctypes.data,
                                          int(sequential), q.ctypes.data if q is not None else None, planes.ctypes.data,
                                          labels.ctypes.data if labels is not None else None,
                                          disp.ctypes.data if disp is not None else None,
                                          depth.ctypes.data if depth is not None else None)
    if rc != 0:
        raise HostError(_lib.cartb200_host_last_error().decode())
    return dict(planes=planes, labels=labels, disparity=disp, depth=depth)


def decode_png(path):
    """The source layer's PNG reader (replaces cv::imread in the KITTI source): returns BGR uint8 [H, W, 3]."""
    w, h = C.c_int(), C.c_int()
    if _lib.cartb200_host_decode_png(str(path).encode(), None, 0, C.byref(w), C.byref(h)) != 0:
        raise HostError(_lib.cartb200_host_last_error().decode())
    out = np.empty((h.value, w.value, 3), np.uint8)
    if _lib.cartb200_host_decode_png(str(path).encode(), out.ctypes.data, out.nbytes, C.byref(w), C.byref(h)) != 0:
        raise HostError(_lib.cartb200_host_last_error().decode())
    return out


def inflate(stream: bytes, size: int, repeat: int = 1):
    """The PNG reader's own zlib-stream decoder: `stream` must decode to exactly `size` bytes (returns them as a uint8
    array) - raises HostError on a malformed / truncated / mis-sized stream or an Adler-32 mismatch."""
    out = np.empty(size, np.uint8)
    if _lib.cartb200_host_inflate(stream, len(stream), out.ctypes.data, size, repeat) != 0:
        raise HostError("inflate failed")
    return out


def open_source(source):
    """Builds a data source from a reference-style source config (dict / JSON text); returns (width, height, Q)."""
    text = source if isinstance(source, str) else json.dumps(source)
    w, h = C.c_int(), C.c_int()
    q = np.zeros(16, np.float32)
    if _lib.cartb200_host_run_source(text.encode(), None, 0, 0, C.byref(w), C.byref(h), q.ctypes.data, None, None, None) < 0:
        raise HostError(_lib.cartb200_host_last_error().decode())
    return w.value, h.value, q.reshape(4, 4)


def run_source(source, modules, max_frames, skip_out_of_scope=False, want_disparity=False, want_depth=False):
    """Runs a module list over the frames of an on-disk source (config/sources/*.json schema), in id order."""
    W, H, Q = open_source(source)
    stext = source if isinstance(source, str) else json.dumps(source)
    mtext = modules if isinstance(modules, str) else json.dumps(modules)
    planes = np.full((max_frames, H, W), 255, np.uint8)
    disp = np.zeros((max_frames, H, W), np.int16) if want_disparity else None
    depth = np.zeros((max_frames, H, W, 3), np.float32) if want_depth else None
    w, h = C.c_int(), C.c_int()
    n = _lib.cartb200_host_run_source(stext.encode(), mtext.encode(), int(skip_out_of_scope), max_frames, C.byref(w), C.byref(h),
                                      None, planes.ctypes.data, disp.ctypes.data if disp is not None else None,
                                      depth.ctypes.data if depth is not None else None)
    if n < 0:
        raise HostError(_lib.cartb200_host_last_error().decode())
    return dict(n=n, Q=Q, planes=planes[:n], disparity=None if disp is None else disp[:n], depth=None if depth is None else depth[:n])
