"""ctypes loader for the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY: import this from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never from
cart_slam_b200/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("sgm.cpp", "stages.cpp", "superpixels.cpp", "tile_capi.cpp", "oracle.h", "tile.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _p(a, t=None):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def gray(bgr):
    bgr = _c(bgr, np.uint8)
    H, W = bgr.shape[:2]
    out = np.empty((H, W), np.uint8)
    lib().orc_gray(_p(bgr), W, H, _p(out))
    return out


def ycrcb(bgr):
    bgr = _c(bgr, np.uint8)
    H, W = bgr.shape[:2]
    out = np.empty((H, W, 3), np.uint8)
    lib().orc_ycrcb(_p(bgr), W, H, _p(out))
    return out


def census(g):
    g = _c(g, np.uint8)
    H, W = g.shape
    out = np.empty((H, W), np.uint32)
    lib().orc_census(_p(g), W, H, _p(out))
    return out


def sgm_dirs(paths):
    out = np.zeros((paths, 2), np.int32)
    assert lib().orc_sgm_dirs(paths, _p(out)) == 0
    return out


def sgm_path(cl, cr, D, min_disp, P1, P2, dx, dy):
    cl, cr = _c(cl, np.uint32), _c(cr, np.uint32)
    H, W = cl.shape
    L = np.empty((H, W, D), np.uint8)
    assert lib().orc_sgm_path(_p(cl), _p(cr), W, H, D, min_disp, P1, P2, dx, dy, _p(L)) == 0
    return L


def sgm_wta(Ls, uniqueness_ratio=12):
    Ls = [_c(L, np.uint8) for L in Ls]
    H, W, D = Ls[0].shape
    arr = (C.c_void_p * len(Ls))(*[L.ctypes.data for L in Ls])
    left = np.empty((H, W), np.uint16)
    right = np.empty((H, W), np.uint16)
    lib().orc_sgm_wta(arr, len(Ls), W, H, D, uniqueness_ratio, _p(left), _p(right))
    return left, right


def median3(img):
    img = _c(img, np.uint16)
    H, W = img.shape
    out = np.empty_like(img)
    lib().orc_median3(_p(img), W, H, _p(out))
    return out


def lr_check_range(left, right, gray_left, min_disp):
    left, right, gray_left = _c(left, np.uint16), _c(right, np.uint16), _c(gray_left, np.uint8)
    H, W = left.shape
    out = np.empty((H, W), np.int16)
    lib().orc_lr_check_range(_p(left), _p(right), _p(gray_left), W, H, min_disp, _p(out))
    return out


def sgm_compute(left_bgr, right_bgr, D, min_disp=4, P1=10, P2=120, uniqueness_ratio=12, paths=4, intermediates=False):
    left_bgr, right_bgr = _c(left_bgr, np.uint8), _c(right_bgr, np.uint8)
    H, W = left_bgr.shape[:2]
    disp = np.empty((H, W), np.int16)
    if intermediates:
        cl = np.empty((H, W), np.uint32)
        cr = np.empty((H, W), np.uint32)
        vol = np.empty((paths, H, W, D), np.uint8)
        lr = np.empty((H, W), np.uint16)
        rr = np.empty((H, W), np.uint16)
    else:
        cl = cr = vol = lr = rr = None
    rc = lib().orc_sgm_compute(_p(left_bgr), _p(right_bgr), W, H, D, min_disp, P1, P2, uniqueness_ratio, paths,
                               _p(disp), _p(cl), _p(cr), _p(vol), _p(lr), _p(rr))
    assert rc == 0, rc
    if intermediates:
        return disp, dict(census_l=cl, census_r=cr, volumes=vol, left_raw=lr, right_raw=rr)
    return disp


def interpolate(disp, radius, iterations, min_disparity, max_disparity, want_mask=False):
    d = _c(disp, np.int16).copy()
    H, W = d.shape
    m = np.zeros((H, W), np.uint8) if want_mask else None
    lib().orc_interpolate(_p(d), W, H, radius, iterations, min_disparity, max_disparity, _p(m))
    return (d, m) if want_mask else d


def derivative(disp, want_mask=False):
    disp = _c(disp, np.int16)
    H, W = disp.shape
    out = np.empty((H, W, 2), np.int16)
    hist = np.zeros((256, 2), np.int32)
    m = np.zeros((H, W, 2), np.uint8) if want_mask else None
    lib().orc_derivative(_p(disp), W, H, _p(out), _p(hist), _p(m))
    return (out, hist, m) if want_mask else (out, hist)


def depth(disp, Q):
    """DepthModule: disparity (x16 fixed point) -> XYZ float32 [H, W, 3] with the 4x4 reprojection matrix Q."""
    disp = _c(disp, np.int16)
    H, W = disp.shape
    q = _c(np.asarray(Q, np.float32).reshape(16), np.float32)
    out = np.empty((H, W, 3), np.float32)
    lib().orc_depth(_p(disp), W, H, _p(q), _p(out))
    return out


def resize_bgr8(img, dw, dh):
    """cv::cuda::resize(..., INTER_LINEAR) on CV_8UC3 as restated in stages.cpp (parity unpinned)."""
    img = _c(img, np.uint8)
    sh, sw, _ = img.shape
    out = np.empty((dh, dw, 3), np.uint8)
    lib().orc_resize_bgr8(_p(img), sw, sh, _p(out), dw, dh)
    return out


def naive_derivative(disp, want_mask=False):
    disp = _c(disp, np.int16)
    H, W = disp.shape
    out = np.empty((H, W), np.int16)
    hist = np.zeros(256, np.int32)
    m = np.zeros((H, W), np.uint8) if want_mask else None
    lib().orc_naive_derivative(_p(disp), W, H, _p(out), _p(hist), _p(m))
    return (out, hist, m) if want_mask else (out, hist)


def classify(deriv, hS, hE, vS, vE, channel_stride=1):
    deriv = _c(deriv, np.int16)
    n = deriv.size // channel_stride
    out = np.empty(n, np.uint8)
    lib().orc_classify.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib().orc_classify(_p(deriv), n, channel_stride, hS, hE, vS, vE, _p(out))
    return out.reshape(deriv.shape[:2])


def sp_planeseg(deriv2, labels, max_label, hS, hE, vS, vE):
    deriv2, labels = _c(deriv2, np.int16), _c(labels, np.uint16)
    H, W = labels.shape
    pu = np.empty((H, W), np.uint8)
    ps = np.empty((H, W), np.uint8)
    rc = lib().orc_sp_planeseg(_p(deriv2), _p(labels), W, H, max_label, hS, hE, vS, vE, _p(pu), _p(ps))
    if rc != 0:
        raise RuntimeError(f"orc_sp_planeseg rc={rc}")
    return pu, ps


def _ptr_array(arrs):
    a = (C.c_void_p * max(1, len(arrs)))()
    for i, x in enumerate(arrs):
        a[i] = x.ctypes.data
    return a


def classify_temporal(deriv, hS, hE, vS, vE, prev_planes, prev_flow, channel_stride=1):
    """prev_planes[k]: planes_unsmoothed of frame id-(k+1) (H, W) u8; prev_flow[k]: optflow of frame id-k (H, W, 2) int16 S10.5."""
    deriv = _c(deriv, np.int16)
    H, W = deriv.shape[:2]
    pp = [_c(a, np.uint8) for a in prev_planes]
    pf = [_c(a, np.int16) for a in prev_flow]
    assert len(pp) == len(pf)
    pu = np.empty((H, W), np.uint8)
    ps = np.empty((H, W), np.uint8)
    f = lib().orc_classify_temporal
    f.argtypes = [C.c_void_p] + [C.c_int] * 8 + [C.c_void_p] * 4
    rc = f(_p(deriv), W, H, channel_stride, hS, hE, vS, vE, len(pp), _ptr_array(pp), _ptr_array(pf), _p(pu), _p(ps))
    if rc != 0:
        raise RuntimeError(f"orc_classify_temporal rc={rc}")
    return pu, ps


def sp_planeseg_temporal(deriv2, labels, max_label, hS, hE, vS, vE, prev_planes, prev_flow):
    deriv2, labels = _c(deriv2, np.int16), _c(labels, np.uint16)
    H, W = labels.shape
    pp = [_c(a, np.uint8) for a in prev_planes]
    pf = [_c(a, np.int16) for a in prev_flow]
    assert len(pp) == len(pf)
    pu = np.empty((H, W), np.uint8)
    ps = np.empty((H, W), np.uint8)
    f = lib().orc_sp_planeseg_temporal
    f.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 8 + [C.c_void_p] * 4
    rc = f(_p(deriv2), _p(labels), W, H, max_label, hS, hE, vS, vE, len(pp), _ptr_array(pp), _ptr_array(pf), _p(pu), _p(ps))
    if rc != 0:
        raise RuntimeError(f"orc_sp_planeseg_temporal rc={rc}")
    return pu, ps


def label_statistics(labels, xyz, n_labels):
    labels, xyz = _c(labels, np.uint16), _c(xyz, np.float32)
    H, W = labels.shape
    cnt = np.empty(n_labels, np.uint32)
    inv = np.empty(n_labels, np.uint32)
    rc = lib().orc_label_statistics(_p(labels), _p(xyz), W, H, n_labels, _p(cnt), _p(inv))
    if rc != 0:
        raise RuntimeError(f"orc_label_statistics rc={rc}")
    return cnt, inv


def region_inliers(labels, xyz, n_labels, planes, threshold):
    labels, xyz = _c(labels, np.uint16), _c(xyz, np.float32)
    planes = _c(np.asarray(planes, np.float64).reshape(-1, 4), np.float64)
    H, W = labels.shape
    out = np.empty((len(planes), n_labels), np.uint32)
    f = lib().orc_region_inliers
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p]
    rc = f(_p(labels), _p(xyz), W, H, n_labels, _p(planes), len(planes), float(threshold), _p(out))
    if rc != 0:
        raise RuntimeError(f"orc_region_inliers rc={rc}")
    return out


def overlay_planes(bgr, planes):
    bgr, planes = _c(bgr, np.uint8), _c(planes, np.uint8)
    H, W = planes.shape
    out = np.empty((H, W, 3), np.uint8)
    rc = lib().orc_overlay_planes(_p(bgr), _p(planes), W, H, _p(out))
    if rc != 0:
        raise RuntimeError(f"orc_overlay_planes rc={rc}")
    return out


def overlay_boundaries(bgr, labels, out=None):
    """`out` (H, W, 3) keeps its last row and column, like the reference's output image."""
    bgr, labels = _c(bgr, np.uint8), _c(labels, np.uint16)
    H, W = labels.shape
    out = np.zeros((H, W, 3), np.uint8) if out is None else _c(out, np.uint8).copy()
    lib().orc_overlay_boundaries(_p(bgr), _p(labels), W, H, _p(out))
    return out


def find_peaks(hist):
    hist = _c(hist, np.int32)
    out = np.zeros((len(hist), 4), np.int32)
    n = lib().orc_find_peaks(_p(hist), len(hist), _p(out), len(hist))
    return out[:n]


def histogram_peak_update(hist, params):
    """params: [hC, vC, hS, hE, vS, vE] -> (updated?, new params)"""
    hist = _c(hist, np.int32)
    p = np.array(params, np.int32)
    r = lib().orc_histogram_peak_update(_p(hist), _p(p))
    return bool(r), [int(v) for v in p]


def block_init(W, H, bw, bh):
    labels = np.empty((H, W), np.uint16)
    n = lib().orc_block_init(W, H, bw, bh, _p(labels))
    return labels, n


def border_map(labels, want_mask=False):
    labels = _c(labels, np.uint16)
    H, W = labels.shape
    b = np.empty((H, W), np.uint8)
    m = np.zeros((H, W), np.uint8) if want_mask else None
    lib().orc_border_map(_p(labels), W, H, _p(b), _p(m))
    return (b, m) if want_mask else b


def sp_relax(labels, max_label, ycrcb_img, deriv2, iterations, direct=0.5, diag=None, w_compact=0.1, progressive=0.0,
             w_disp=1.0, w_image=1.5):
    if diag is None:
        diag = direct / np.sqrt(2)
    lab = _c(labels, np.uint16).copy()
    H, W = lab.shape
    yc = _c(ycrcb_img, np.uint8)
    dv = _c(deriv2, np.int16) if deriv2 is not None else None
    bc = np.zeros(max(1, iterations), np.int32)
    mv = np.zeros(max(1, iterations), np.int32)
    f = lib().orc_sp_relax
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_double] * 6 + [C.c_void_p, C.c_void_p]
    rc = f(_p(lab), W, H, max_label, _p(yc), _p(dv), iterations, direct, diag, w_compact, progressive, w_disp, w_image, _p(bc), _p(mv))
    if rc != 0:
        raise RuntimeError(f"orc_sp_relax rc={rc}")
    return lab, bc[:iterations], mv[:iterations]


def tile_i32(img, bx, by, bdx, bdy, XB, YB, y_pad, x_pad, interp, alloc_elems=None, undef=-1):
    img = _c(img, np.int32)
    H, W = img.shape
    S, R = XB * bdx + 2 * x_pad, YB * bdy + 2 * y_pad
    if alloc_elems is None:
        alloc_elems = S * R
    out = np.empty((R, S), np.int32)
    d = np.empty((R, S), np.uint8)
    f = lib().orc_tile_i32
    f.argtypes = [C.c_void_p] + [C.c_int] * 11 + [C.c_long, C.c_int, C.c_void_p, C.c_void_p]
    f(_p(img), W, H, bx, by, bdx, bdy, XB, YB, y_pad, x_pad, int(interp), alloc_elems, undef, _p(out), _p(d))
    return out, d
