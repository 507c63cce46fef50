"""cart_slam_b200 - Python binding (ctypes) of libcartb200, the B200-native disparity -> planeseg path.

PyTorch is used for device memory and streams only; every computation happens in the hand-written
sm_100a kernels behind the C ABI declared in include/cartb200.h.  There is no CPU fallback: the first
call into the package without the compiled library, or creating a Context without a GPU, fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("CARTB200_LIB") or os.path.join(_HERE, "libcartb200.so")  # CARTB200_LIB: experiments with another build

OK, E_ARG, E_SHAPE, E_CUDA, E_NOMEM, E_UNSUPPORTED = 0, -1, -2, -3, -4, -5
DISPARITY_INVALID = -32768
PLANE_HORIZONTAL, PLANE_VERTICAL, PLANE_UNKNOWN = 0, 1, 2

# data keys of the reference's module hand-off
KEY_DISPARITY = "disparity"
KEY_DISPARITY_DERIVATIVE = "disparity_derivative"
KEY_DISPARITY_DERIVATIVE_HISTOGRAM = "disparity_derivative_histogram"
KEY_SUPERPIXELS = "superpixels"
KEY_SUPERPIXELS_MAX_LABEL = "superpixels_max_label"
KEY_PLANES = "planes"
KEY_PLANES_UNSMOOTHED = "planes_unsmoothed"


class CartB200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"cartb200 error {code}: {message}")
        self.code = code


class _CConfig(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int), ("max_batch", C.c_int), ("enable_sgm", C.c_int),
        ("min_disparity", C.c_int), ("num_disparities", C.c_int), ("p1", C.c_int), ("p2", C.c_int),
        ("uniqueness_ratio", C.c_int), ("paths", C.c_int), ("smoothing_radius", C.c_int),
        ("smoothing_iterations", C.c_int), ("enable_superpixels", C.c_int), ("sp_block_size", C.c_int),
        ("sp_direct_clique_cost", C.c_double), ("sp_diagonal_clique_cost", C.c_double),
        ("sp_compactness_weight", C.c_double), ("sp_progressive_compactness_cost", C.c_double),
        ("sp_image_weight", C.c_double), ("sp_disparity_weight", C.c_double), ("sp_exact", C.c_int),
    ]


class _CSeqOpts(C.Structure):
    _fields_ = [
        ("pipeline", C.c_int), ("provider", C.c_int), ("static_params", C.c_int * 4),
        ("update_interval", C.c_int), ("reset_interval", C.c_int), ("sp_initial_iterations", C.c_int),
        ("sp_iterations", C.c_int), ("sp_reset_iterations", C.c_int), ("start_id", C.c_int),
    ]


class _CTemporalRef(C.Structure):  # cartb200_temporal_ref
    _fields_ = [("planes_unsmoothed", C.c_void_p), ("planes_pitch", C.c_size_t), ("optflow", C.c_void_p),
                ("optflow_pitch", C.c_size_t)]


def _load():
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "cart_slam_b200 has no CPU fallback."
        )
    lib = C.CDLL(_LIB_PATH)
    lib.cartb200_version.restype = C.c_char_p
    lib.cartb200_last_error.restype = C.c_char_p
    lib.cartb200_last_error.argtypes = [C.c_void_p]
    lib.cartb200_launch_count.restype = C.c_longlong
    lib.cartb200_launch_count.argtypes = [C.c_void_p]
    lib.cartb200_scratch_bytes.restype = C.c_size_t
    lib.cartb200_scratch_bytes.argtypes = [C.c_void_p]
    lib.cartb200_create.argtypes = [C.POINTER(_CConfig), C.POINTER(C.c_void_p)]
    lib.cartb200_destroy.argtypes = [C.c_void_p]
    lib.cartb200_default_config.argtypes = [C.POINTER(_CConfig), C.c_int, C.c_int]
    lib.cartb200_default_sequence_opts.argtypes = [C.POINTER(_CSeqOpts)]
    vp, sz, i = C.c_void_p, C.c_size_t, C.c_int
    lib.cartb200_disparity.argtypes = [vp, i, vp, vp, sz, sz, vp, sz, sz, vp]
    lib.cartb200_sgm_gray_census.argtypes = [vp, i, vp, vp, sz, sz, vp]
    lib.cartb200_sgm_aggregate.argtypes = [vp, i, vp]
    lib.cartb200_sgm_aggregate_path.argtypes = [vp, i, i, vp]
    lib.cartb200_sgm_wta_post.argtypes = [vp, i, vp, sz, sz, vp]
    lib.cartb200_interpolate.argtypes = [vp, i, vp, sz, sz, i, i, i, i, vp]
    lib.cartb200_sgm_intermediate.argtypes = [vp, i, C.POINTER(vp), C.POINTER(sz), C.POINTER(sz)]
    lib.cartb200_last_create_error.restype = C.c_char_p
    lib.cartb200_derivative.argtypes = [vp, i, vp, sz, sz, vp, sz, sz, vp, vp]
    lib.cartb200_depth.argtypes = [vp, i, vp, sz, sz, vp, vp, sz, sz, vp]
    lib.cartb200_naive_derivative.argtypes = [vp, i, vp, sz, sz, vp, sz, sz, vp, vp]
    lib.cartb200_classify.argtypes = [vp, i, vp, sz, sz, i, i, vp, vp, sz, sz, vp]
    lib.cartb200_superpixels_reset.argtypes = [vp, i, vp, C.POINTER(i), vp]
    lib.cartb200_superpixels_relax.argtypes = [vp, i, vp, i, vp, sz, sz, vp, sz, sz, vp, sz, sz, vp]
    lib.cartb200_superpixels_set_labels.argtypes = [vp, i, vp, sz, vp]
    lib.cartb200_superpixels_border_map.argtypes = [vp, vp, sz, vp, sz, vp]
    lib.cartb200_sp_planeseg.argtypes = [vp, i, vp, sz, sz, vp, sz, sz, i, vp, vp, vp, sz, sz, vp]
    lib.cartb200_histogram_peak_update.argtypes = [vp, vp]
    lib.cartb200_classify_temporal.argtypes = [vp, vp, sz, i, i, vp, i, vp, vp, vp, sz, vp]
    lib.cartb200_sp_planeseg_temporal.argtypes = [vp, vp, sz, vp, sz, i, vp, i, vp, vp, vp, sz, vp]
    lib.cartb200_overlay_planes.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp]
    lib.cartb200_overlay_superpixel_boundaries.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp]
    lib.cartb200_label_statistics.argtypes = [vp, vp, sz, vp, sz, i, vp, vp, vp]
    lib.cartb200_region_inliers.argtypes = [vp, vp, sz, vp, sz, i, vp, i, C.c_double, vp, vp]
    lib.cartb200_run_sequence_host.argtypes = [vp, C.POINTER(_CSeqOpts), i, vp, vp, vp, vp]
    lib.cartb200_run_sequence_device.argtypes = [vp, C.POINTER(_CSeqOpts), i, vp, vp, vp, vp, vp]
    lib.cartb200_resize_bgr8.argtypes = [vp, sz, i, i, vp, sz, i, i, vp]
    lib.cartb200_run_sequence_phase1_device.argtypes = [vp, C.POINTER(_CSeqOpts), i, vp, vp, vp, vp, vp]
    lib.cartb200_run_sequence_phase2_device.argtypes = [vp, C.POINTER(_CSeqOpts), i, vp, vp, vp]
    lib.cartb200_sequence_parameters.argtypes = [C.POINTER(_CSeqOpts), i, vp, vp]
    lib.cartb200_run_sequence_phase1_host.argtypes = [vp, C.POINTER(_CSeqOpts), i, vp, vp, vp, vp]
    lib.cartb200_run_sequence_phase2_host.argtypes = [vp, C.POINTER(_CSeqOpts), i, vp, vp]
    lib.cartb200_debug_ref_tile_i32.restype = C.c_int32
    lib.cartb200_debug_ref_tile_i32.argtypes = [vp] + [i] * 11 + [C.c_long, C.c_int32, i, i]
    lib.cartb200_debug_median9.restype = C.c_uint32
    lib.cartb200_debug_median9.argtypes = [vp]
    return lib


class _LazyLib:
    """libcartb200.so is mapped on first use, not at import: `import cart_slam_b200.synth` (input generation, also used
    by the CPU reference arm of bench.py) must not load the CUDA library."""

    _lib = None

    def __getattr__(self, name):
        if _LazyLib._lib is None:
            _LazyLib._lib = _load()
        return getattr(_LazyLib._lib, name)


_lib = _LazyLib()

EXPORTED_SYMBOLS = [
    "cartb200_default_config", "cartb200_create", "cartb200_last_create_error", "cartb200_destroy", "cartb200_last_error", "cartb200_version",
    "cartb200_launch_count", "cartb200_scratch_bytes", "cartb200_disparity", "cartb200_sgm_gray_census",
    "cartb200_sgm_aggregate", "cartb200_sgm_aggregate_path", "cartb200_sgm_wta_post", "cartb200_interpolate", "cartb200_sgm_intermediate",
    "cartb200_derivative", "cartb200_depth", "cartb200_naive_derivative", "cartb200_classify", "cartb200_superpixels_reset",
    "cartb200_superpixels_relax", "cartb200_superpixels_set_labels", "cartb200_superpixels_border_map",
    "cartb200_sp_planeseg", "cartb200_histogram_peak_update", "cartb200_default_sequence_opts",
    "cartb200_run_sequence_host", "cartb200_run_sequence_device", "cartb200_run_sequence_phase1_device",
    "cartb200_run_sequence_phase2_device", "cartb200_run_sequence_phase1_host", "cartb200_run_sequence_phase2_host",
    "cartb200_sequence_parameters", "cartb200_resize_bgr8", "cartb200_debug_ref_tile_i32",
]


def resize_bgr8(src, dw: int, dh: int):
    """cv::cuda::resize(..., INTER_LINEAR) replacement for CV_8UC3 device images: src = uint8 CUDA tensor [H, W, 3]."""
    import torch
    assert src.is_cuda and src.dtype == torch.uint8 and src.dim() == 3 and src.shape[2] == 3 and src.is_contiguous()
    out = torch.empty((dh, dw, 3), dtype=torch.uint8, device=src.device)
    rc = _lib.cartb200_resize_bgr8(src.data_ptr(), src.shape[1] * 3, src.shape[1], src.shape[0], out.data_ptr(), dw * 3, dw, dh,
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise CartB200Error(rc, "resize_bgr8 failed")
    return out


def sequence_parameters(opts: "SequenceOptions", hist: np.ndarray) -> np.ndarray:
    """The reference's parameter bookkeeping over a whole sequence (CPU, no GPU needed): hist = [n, 256] int32 per-frame
    histograms in id order starting at opts.start_id; returns [n, 4] int32 {hStart, hEnd, vStart, vEnd} per frame."""
    hist = np.ascontiguousarray(hist, dtype=np.int32)
    assert hist.ndim == 2 and hist.shape[1] == 256
    out = np.zeros((hist.shape[0], 4), np.int32)
    o = opts.to_c()
    rc = _lib.cartb200_sequence_parameters(C.byref(o), hist.shape[0], hist.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise ValueError(f"cartb200_sequence_parameters: bad arguments ({rc})")
    return out


def version() -> str:
    return _lib.cartb200_version().decode()


def library_path() -> str:
    return _LIB_PATH


@dataclass
class Config:
    """Mirror of cartb200_config; defaults = the reference's JSON defaults (cartconfig.cpp:121-152)."""
    width: int
    height: int
    max_batch: int = 1
    enable_sgm: bool = True
    min_disparity: int = 4
    num_disparities: int = 256
    p1: int = 10
    p2: int = 120
    uniqueness_ratio: int = 12
    paths: int = 4
    smoothing_radius: int = -1
    smoothing_iterations: int = 5
    enable_superpixels: bool = True
    sp_block_size: int = 12
    sp_direct_clique_cost: float = 0.5
    sp_diagonal_clique_cost: Optional[float] = None
    sp_compactness_weight: float = 0.1
    sp_progressive_compactness_cost: float = 0.0
    sp_image_weight: float = 1.5
    sp_disparity_weight: float = 1.0
    sp_exact: bool = True  # kept for ABI compatibility, ignored: label costs are always in the reference's operation order

    def to_c(self) -> _CConfig:
        c = _CConfig()
        _lib.cartb200_default_config(C.byref(c), self.width, self.height)
        for name, _ in _CConfig._fields_:
            v = getattr(self, name)
            if name == "sp_diagonal_clique_cost" and v is None:
                v = self.sp_direct_clique_cost / np.sqrt(2.0)
            setattr(c, name, int(v) if isinstance(getattr(c, name), int) else float(v))
        return c


@dataclass
class SequenceOptions:
    pipeline: int = 0            # 0 naive (kitti-naive-segmentation.json), 1 superpixel (kitti-planeseg.json)
    provider: int = 1            # 0 static, 1 histogram_peak
    static_params: Sequence[int] = field(default_factory=lambda: [1, 30, -3, 1])
    update_interval: int = 30
    reset_interval: int = 10
    sp_initial_iterations: int = 18
    sp_iterations: int = 6
    sp_reset_iterations: int = 64
    start_id: int = 1

    def to_c(self) -> _CSeqOpts:
        o = _CSeqOpts()
        _lib.cartb200_default_sequence_opts(C.byref(o))
        o.pipeline, o.provider = self.pipeline, self.provider
        for k in range(4):
            o.static_params[k] = int(self.static_params[k])
        o.update_interval, o.reset_interval = self.update_interval, self.reset_interval
        o.sp_initial_iterations, o.sp_iterations = self.sp_initial_iterations, self.sp_iterations
        o.sp_reset_iterations, o.start_id = self.sp_reset_iterations, self.start_id
        return o


def histogram_peak_update(hist256, params):
    """HistogramPeakPlaneParameterProvider::updatePlaneParameters. params = [hC, vC, hS, hE, vS, vE]."""
    h = np.ascontiguousarray(hist256, dtype=np.int32)
    p = np.array(params, dtype=np.int32)
    assert h.size == 256 and p.size == 6
    r = _lib.cartb200_histogram_peak_update(h.ctypes.data, p.ctypes.data)
    if r < 0:
        raise CartB200Error(r, "histogram_peak_update")
    return bool(r), [int(v) for v in p]


def debug_ref_tile_i32(img, bx, by, bdx, bdy, XB, YB, y_pad, x_pad, interp, lx, ly, alloc_elems=None, undef=-1):
    img = np.ascontiguousarray(img, dtype=np.int32)
    H, W = img.shape
    if alloc_elems is None:
        alloc_elems = (XB * bdx + 2 * x_pad) * (YB * bdy + 2 * y_pad)
    return int(_lib.cartb200_debug_ref_tile_i32(img.ctypes.data, W, H, bx, by, bdx, bdy, XB, YB, y_pad, x_pad,
                                                int(interp), alloc_elems, undef, lx, ly))


def debug_median9(v9) -> int:
    v = np.ascontiguousarray(v9, dtype=np.uint16)
    assert v.size == 9
    return int(_lib.cartb200_debug_median9(v.ctypes.data))


class _DevView:
    def __init__(self, ptr, shape, typestr, strides):
        self.__cuda_array_interface__ = {"data": (ptr, False), "shape": shape, "typestr": typestr,
                                         "strides": strides, "version": 2}


class Context:
    """One cartb200 context (scratch sized for `max_batch` frames). Not thread-safe."""

    def __init__(self, cfg: Config):
        import torch

        if not torch.cuda.is_available():
            raise CartB200Error(E_CUDA, "no CUDA device: the cartb200 path has no CPU fallback")
        self.torch = torch
        self.cfg = cfg
        self._h = C.c_void_p()
        c = cfg.to_c()
        rc = _lib.cartb200_create(C.byref(c), C.byref(self._h))
        if rc != OK:
            raise CartB200Error(rc, _lib.cartb200_last_create_error().decode() or "cartb200_create failed")
        self.W, self.H, self.D, self.B = cfg.width, cfg.height, cfg.num_disparities, cfg.max_batch
        self.max_label = -(-cfg.width // cfg.sp_block_size) * -(-cfg.height // cfg.sp_block_size)

    def close(self):
        if self._h:
            _lib.cartb200_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != OK:
            raise CartB200Error(rc, _lib.cartb200_last_error(self._h).decode())

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def _img(self, t, dtype, last=None):
        torch = self.torch
        assert t.is_cuda and t.dtype == dtype and t.is_contiguous(), (t.dtype, t.is_cuda, t.is_contiguous())
        assert t.shape[1] == self.H and t.shape[2] == self.W, t.shape
        if last is not None:
            assert t.dim() == 4 and t.shape[3] == last
        n = t.shape[0]
        pitch = t.stride(1) * t.element_size()
        fstride = t.stride(0) * t.element_size()
        return n, C.c_void_p(t.data_ptr()), pitch, fstride

    def launch_count(self) -> int:
        return int(_lib.cartb200_launch_count(self._h))

    def scratch_bytes(self) -> int:
        return int(_lib.cartb200_scratch_bytes(self._h))

    # -- disparity -------------------------------------------------------------------------------
    def disparity(self, left, right):
        torch = self.torch
        n, lp, pitch, fs = self._img(left, torch.uint8, 3)
        n2, rp, pitch2, fs2 = self._img(right, torch.uint8, 3)
        assert (n, pitch, fs) == (n2, pitch2, fs2)
        out = torch.empty((n, self.H, self.W), dtype=torch.int16, device=left.device)
        self._check(_lib.cartb200_disparity(self._h, n, lp, rp, pitch, fs, out.data_ptr(), self.W * 2,
                                            self.W * self.H * 2, self._stream()))
        return out

    def sgm_gray_census(self, left, right):
        torch = self.torch
        n, lp, pitch, fs = self._img(left, torch.uint8, 3)
        _, rp, _, _ = self._img(right, torch.uint8, 3)
        self._check(_lib.cartb200_sgm_gray_census(self._h, n, lp, rp, pitch, fs, self._stream()))
        return n

    def sgm_aggregate(self, n):
        self._check(_lib.cartb200_sgm_aggregate(self._h, n, self._stream()))

    def sgm_aggregate_path(self, n, path):
        self._check(_lib.cartb200_sgm_aggregate_path(self._h, n, path, self._stream()))

    def sgm_wta_post(self, n):
        torch = self.torch
        out = torch.empty((n, self.H, self.W), dtype=torch.int16, device="cuda")
        self._check(_lib.cartb200_sgm_wta_post(self._h, n, out.data_ptr(), self.W * 2, self.W * self.H * 2, self._stream()))
        return out

    def sgm_intermediate(self, which: int, n: int):
        """Copy of an intermediate of the last SGM call as a torch tensor [n, H, W(, D)]."""
        torch = self.torch
        ptr, pitch, fs = C.c_void_p(), C.c_size_t(), C.c_size_t()
        self._check(_lib.cartb200_sgm_intermediate(self._h, which, C.byref(ptr), C.byref(pitch), C.byref(fs)))
        if which >= 10:
            v = _DevView(ptr.value, (n, self.H, self.W, self.D), "|u1", (fs.value, pitch.value, self.D, 1))
        else:
            ts, es = {0: ("<u4", 4), 1: ("<u4", 4), 2: ("|u1", 1), 3: ("<u2", 2), 4: ("<u2", 2)}[which]
            v = _DevView(ptr.value, (n, self.H, self.W), ts, (fs.value, pitch.value, es))
        t = torch.as_tensor(v, device="cuda")
        return t.clone()

    def interpolate(self, disp, radius, iterations, min_disparity, max_disparity):
        torch = self.torch
        n, p, pitch, fs = self._img(disp, torch.int16)
        self._check(_lib.cartb200_interpolate(self._h, n, p, pitch, fs, radius, iterations, min_disparity,
                                              max_disparity, self._stream()))
        return disp

    # -- derivative / planeseg -------------------------------------------------------------------
    def derivative(self, disp):
        torch = self.torch
        n, p, pitch, fs = self._img(disp, torch.int16)
        deriv = torch.empty((n, self.H, self.W, 2), dtype=torch.int16, device=disp.device)
        hist = torch.empty((n, 256, 2), dtype=torch.int32, device=disp.device)
        self._check(_lib.cartb200_derivative(self._h, n, p, pitch, fs, deriv.data_ptr(), self.W * 4,
                                             self.W * self.H * 4, hist.data_ptr(), self._stream()))
        return deriv, hist

    def depth(self, disp, Q):
        """DepthModule: disparity [n,H,W] int16 (x16) -> XYZ float32 [n,H,W,3]; Q = 4x4 reprojection matrix."""
        torch = self.torch
        n, p, pitch, fs = self._img(disp, torch.int16)
        q = np.ascontiguousarray(np.asarray(Q, np.float32).reshape(16))
        xyz = torch.empty((n, self.H, self.W, 3), dtype=torch.float32, device=disp.device)
        self._check(_lib.cartb200_depth(self._h, n, p, pitch, fs, q.ctypes.data, xyz.data_ptr(), self.W * 12,
                                        self.W * self.H * 12, self._stream()))
        return xyz

    def naive_derivative(self, disp):
        torch = self.torch
        n, p, pitch, fs = self._img(disp, torch.int16)
        deriv = torch.empty((n, self.H, self.W), dtype=torch.int16, device=disp.device)
        hist = torch.empty((n, 256), dtype=torch.int32, device=disp.device)
        self._check(_lib.cartb200_naive_derivative(self._h, n, p, pitch, fs, deriv.data_ptr(), self.W * 2,
                                                   self.W * self.H * 2, hist.data_ptr(), self._stream()))
        return deriv, hist

    def _params(self, params, n):
        p = np.ascontiguousarray(params, dtype=np.int32).reshape(-1, 4)
        if p.shape[0] == 1 and n > 1:
            p = np.repeat(p, n, axis=0)
        assert p.shape[0] == n
        return np.ascontiguousarray(p)

    def classify(self, deriv, params):
        """deriv: [n,H,W] (naive) or [n,H,W,2] (channel 0 is used). params: [n][hS,hE,vS,vE]."""
        torch = self.torch
        channels = 1 if deriv.dim() == 3 else deriv.shape[3]
        n, p, pitch, fs = self._img(deriv, torch.int16)
        pr = self._params(params, n)
        planes = torch.empty((n, self.H, self.W), dtype=torch.uint8, device=deriv.device)
        self._check(_lib.cartb200_classify(self._h, n, p, pitch, fs, channels, 0, pr.ctypes.data, planes.data_ptr(),
                                           self.W, self.W * self.H, self._stream()))
        torch.cuda.current_stream().synchronize()  # params were read from pageable host memory
        return planes

    def sp_planeseg(self, deriv, labels, params, max_label=None):
        torch = self.torch
        n, dp, dpitch, dfs = self._img(deriv, torch.int16, 2)
        n2, lp, lpitch, lfs = self._img(labels, torch.uint16)
        assert n == n2
        pr = self._params(params, n)
        unsm = torch.empty((n, self.H, self.W), dtype=torch.uint8, device=deriv.device)
        planes = torch.empty_like(unsm)
        self._check(_lib.cartb200_sp_planeseg(self._h, n, dp, dpitch, dfs, lp, lpitch, lfs,
                                              self.max_label if max_label is None else max_label, pr.ctypes.data,
                                              unsm.data_ptr(), planes.data_ptr(), self.W, self.W * self.H, self._stream()))
        torch.cuda.current_stream().synchronize()
        return unsm, planes

    # -- temporal smoothing vote --------------------------------------------------
    def _temporal_refs(self, prev_planes, prev_flow):
        torch = self.torch
        assert len(prev_planes) == len(prev_flow)
        arr = (_CTemporalRef * max(1, len(prev_planes)))()
        for k, (pl, fl) in enumerate(zip(prev_planes, prev_flow)):
            assert pl.is_cuda and pl.dtype == torch.uint8 and pl.shape == (self.H, self.W) and pl.stride(1) == 1
            assert fl.is_cuda and fl.dtype == torch.int16 and fl.shape == (self.H, self.W, 2) and fl.stride(2) == 1 and fl.stride(1) == 2
            arr[k] = _CTemporalRef(pl.data_ptr(), pl.stride(0), fl.data_ptr(), fl.stride(0) * 2)
        return arr

    def classify_temporal(self, deriv, params, prev_planes, prev_flow):
        """One frame. deriv: [H,W] or [H,W,2] int16; prev_planes[k]: planes_unsmoothed of frame id-(k+1) [H,W] u8;
        prev_flow[k]: optflow of frame id-k [H,W,2] int16 (S10.5).  Returns (planes_unsmoothed, planes)."""
        torch = self.torch
        channels = 1 if deriv.dim() == 2 else deriv.shape[2]
        _, p, pitch, _ = self._img(deriv[None], torch.int16)
        pr = self._params(params, 1)
        refs = self._temporal_refs(prev_planes, prev_flow)
        unsm = torch.empty((self.H, self.W), dtype=torch.uint8, device=deriv.device)
        sm = torch.empty_like(unsm)
        self._check(_lib.cartb200_classify_temporal(self._h, p, pitch, channels, 0, pr.ctypes.data, len(prev_planes), refs,
                                                    unsm.data_ptr(), sm.data_ptr(), self.W, self._stream()))
        return unsm, sm

    def sp_planeseg_temporal(self, deriv, labels, params, prev_planes, prev_flow, max_label=None):
        torch = self.torch
        _, dp, dpitch, _ = self._img(deriv[None], torch.int16, 2)
        _, lp, lpitch, _ = self._img(labels[None], torch.uint16)
        pr = self._params(params, 1)
        refs = self._temporal_refs(prev_planes, prev_flow)
        unsm = torch.empty((self.H, self.W), dtype=torch.uint8, device=deriv.device)
        planes = torch.empty_like(unsm)
        self._check(_lib.cartb200_sp_planeseg_temporal(self._h, dp, dpitch, lp, lpitch,
                                                       self.max_label if max_label is None else max_label, pr.ctypes.data,
                                                       len(prev_planes), refs, unsm.data_ptr(), planes.data_ptr(), self.W,
                                                       self._stream()))
        return unsm, planes

    # -- superpixel consumers of the plane fit -----------------------------------
    def label_statistics(self, labels, xyz, n_labels):
        """countPixels: labels [H,W] u16, xyz [H,W,3] f32 -> (pixel_count, pixel_count_invalid) uint32 [n_labels] (as int32 tensors)."""
        torch = self.torch
        _, lp, lpitch, _ = self._img(labels[None], torch.uint16)
        _, xp, xpitch, _ = self._img(xyz[None], torch.float32, 3)
        cnt = torch.empty(n_labels, dtype=torch.int32, device=labels.device)
        inv = torch.empty_like(cnt)
        self._check(_lib.cartb200_label_statistics(self._h, lp, lpitch, xp, xpitch, n_labels, cnt.data_ptr(), inv.data_ptr(),
                                                   self._stream()))
        return cnt, inv

    def region_inliers(self, labels, xyz, n_labels, planes, threshold):
        """calculateRegionDistance: planes [n_planes][a,b,c,d] (host) -> inliers int32 [n_planes, n_labels]."""
        torch = self.torch
        _, lp, lpitch, _ = self._img(labels[None], torch.uint16)
        _, xp, xpitch, _ = self._img(xyz[None], torch.float32, 3)
        pl = np.ascontiguousarray(np.asarray(planes, np.float64).reshape(-1, 4))
        out = torch.empty((pl.shape[0], n_labels), dtype=torch.int32, device=labels.device)
        self._check(_lib.cartb200_region_inliers(self._h, lp, lpitch, xp, xpitch, n_labels, pl.ctypes.data, pl.shape[0],
                                                 float(threshold), out.data_ptr(), self._stream()))
        return out

    def overlay_planes(self, image_bgr, planes):
        """overlayPlanes: image [H,W,3] u8 + planes [H,W] u8 -> blended BGR image."""
        torch = self.torch
        _, ip, ipitch, _ = self._img(image_bgr[None], torch.uint8, 3)
        _, pp, ppitch, _ = self._img(planes[None], torch.uint8)
        out = torch.empty_like(image_bgr)
        self._check(_lib.cartb200_overlay_planes(self._h, ip, ipitch, pp, ppitch, out.data_ptr(), self.W * 3, self._stream()))
        return out

    def overlay_superpixel_boundaries(self, image_bgr, labels, out=None):
        """overlayBoundaryVisualization; `out` keeps its last row / column (zeros when not given)."""
        torch = self.torch
        _, ip, ipitch, _ = self._img(image_bgr[None], torch.uint8, 3)
        _, lp, lpitch, _ = self._img(labels[None], torch.uint16)
        out = torch.zeros_like(image_bgr) if out is None else out
        self._check(_lib.cartb200_overlay_superpixel_boundaries(self._h, ip, ipitch, lp, lpitch, out.data_ptr(), self.W * 3,
                                                                self._stream()))
        return out

    # -- superpixels -----------------------------------------------------------------------------
    def _slots(self, slots, n):
        if slots is None:
            return None, None
        a = np.ascontiguousarray(slots, dtype=np.int32)
        assert a.size == n
        return a, C.c_void_p(a.ctypes.data)

    def superpixels_reset(self, n=1, slots=None) -> int:
        keep, sp = self._slots(slots, n)
        ml = C.c_int()
        self._check(_lib.cartb200_superpixels_reset(self._h, n, sp, C.byref(ml), self._stream()))
        self.torch.cuda.current_stream().synchronize()
        return ml.value

    def superpixels_relax(self, left, deriv, iterations, slots=None):
        torch = self.torch
        n, lp, pitch, fs = self._img(left, torch.uint8, 3)
        keep, sp = self._slots(slots, n)
        if deriv is not None:
            n2, dp, dpitch, dfs = self._img(deriv, torch.int16, 2)
            assert n2 == n
        else:
            dp, dpitch, dfs = None, 0, 0
        out = torch.empty((n, self.H, self.W), dtype=torch.uint16, device=left.device)
        self._check(_lib.cartb200_superpixels_relax(self._h, n, sp, iterations, lp, pitch, fs, dp, dpitch, dfs,
                                                    out.data_ptr(), self.W * 2, self.W * self.H * 2, self._stream()))
        torch.cuda.current_stream().synchronize()
        return out

    def superpixels_set_labels(self, slot, labels):
        torch = self.torch
        assert labels.is_cuda and labels.dtype == torch.uint16 and labels.shape == (self.H, self.W) and labels.is_contiguous()
        self._check(_lib.cartb200_superpixels_set_labels(self._h, slot, labels.data_ptr(), self.W * 2, self._stream()))

    def superpixels_border_map(self, labels):
        torch = self.torch
        assert labels.is_cuda and labels.dtype == torch.uint16 and labels.shape == (self.H, self.W) and labels.is_contiguous()
        out = torch.empty((self.H, self.W), dtype=torch.uint8, device=labels.device)
        self._check(_lib.cartb200_superpixels_border_map(self._h, labels.data_ptr(), self.W * 2, out.data_ptr(), self.W,
                                                         self._stream()))
        return out

    # -- whole sequence --------------------------------------------------------------------------
    def run_sequence_host(self, opts: SequenceOptions, left, right, want_disparity=False, planes_out=None,
                          disparity_out=None):
        """left/right: host uint8 [n,H,W,3] (numpy arrays or pinned CPU torch tensors). Returns numpy planes."""
        def ptr(a):
            if isinstance(a, np.ndarray):
                assert a.flags.c_contiguous
                return a.ctypes.data
            return a.data_ptr()
        n = left.shape[0]
        assert tuple(left.shape) == (n, self.H, self.W, 3) and tuple(right.shape) == tuple(left.shape)
        if planes_out is None:
            planes_out = np.empty((n, self.H, self.W), np.uint8)
        if want_disparity and disparity_out is None:
            disparity_out = np.empty((n, self.H, self.W), np.int16)
        o = opts.to_c()
        self._check(_lib.cartb200_run_sequence_host(self._h, C.byref(o), n, ptr(left), ptr(right), ptr(planes_out),
                                                    ptr(disparity_out) if disparity_out is not None else None))
        return (planes_out, disparity_out) if want_disparity else planes_out

    @staticmethod
    def _hptr(a):
        if isinstance(a, np.ndarray):
            assert a.flags.c_contiguous
            return a.ctypes.data
        return a.data_ptr()

    def run_sequence_phase1_host(self, opts: SequenceOptions, left, right) -> np.ndarray:
        """Phase 1 with HOST image buffers (numpy or pinned CPU tensors); returns the histograms [n, 256]."""
        n = left.shape[0]
        assert tuple(left.shape) == (n, self.H, self.W, 3) and tuple(right.shape) == tuple(left.shape)
        hist = np.zeros((n, 256), np.int32)
        o = opts.to_c()
        self._check(_lib.cartb200_run_sequence_phase1_host(self._h, C.byref(o), n, self._hptr(left), self._hptr(right),
                                                           hist.ctypes.data, None))
        return hist

    def run_sequence_phase2_host(self, opts: SequenceOptions, params: np.ndarray, planes_out=None):
        """Phase 2 with a HOST plane buffer."""
        params = np.ascontiguousarray(params, dtype=np.int32)
        n = params.shape[0]
        if planes_out is None:
            planes_out = np.empty((n, self.H, self.W), np.uint8)
        o = opts.to_c()
        self._check(_lib.cartb200_run_sequence_phase2_host(self._h, C.byref(o), n, params.ctypes.data, self._hptr(planes_out)))
        return planes_out

    def run_sequence_phase1(self, opts: SequenceOptions, left, right) -> np.ndarray:
        """Phase 1 of the sharded runner on device buffers; returns the per-frame histograms [n, 256] int32 (host)."""
        torch = self.torch
        n, lp, pitch, fs = self._img(left, torch.uint8, 3)
        _, rp, _, _ = self._img(right, torch.uint8, 3)
        assert pitch == self.W * 3 and fs == self.W * self.H * 3, "device sequence buffers must be tightly packed"
        hist = np.zeros((n, 256), np.int32)
        o = opts.to_c()
        self._check(_lib.cartb200_run_sequence_phase1_device(self._h, C.byref(o), n, lp, rp, hist.ctypes.data, None, self._stream()))
        return hist

    def run_sequence_phase2(self, opts: SequenceOptions, params: np.ndarray, planes_out=None, device=None):
        """Phase 2: params = [n, 4] int32 (host) for the frames of the preceding phase 1; returns device planes."""
        torch = self.torch
        params = np.ascontiguousarray(params, dtype=np.int32)
        n = params.shape[0]
        assert params.shape == (n, 4)
        if planes_out is None:
            planes_out = torch.empty((n, self.H, self.W), dtype=torch.uint8, device=device or "cuda")
        o = opts.to_c()
        self._check(_lib.cartb200_run_sequence_phase2_device(self._h, C.byref(o), n, params.ctypes.data, planes_out.data_ptr(),
                                                             self._stream()))
        return planes_out

    def run_sequence_device(self, opts: SequenceOptions, left, right, want_disparity=False, planes_out=None):
        torch = self.torch
        n, lp, pitch, fs = self._img(left, torch.uint8, 3)
        _, rp, _, _ = self._img(right, torch.uint8, 3)
        assert pitch == self.W * 3 and fs == self.W * self.H * 3, "device sequence buffers must be tightly packed"
        if planes_out is None:
            planes_out = torch.empty((n, self.H, self.W), dtype=torch.uint8, device=left.device)
        disp = torch.empty((n, self.H, self.W), dtype=torch.int16, device=left.device) if want_disparity else None
        o = opts.to_c()
        self._check(_lib.cartb200_run_sequence_device(self._h, C.byref(o), n, lp, rp, planes_out.data_ptr(),
                                                      disp.data_ptr() if disp is not None else None, self._stream()))
        return (planes_out, disp) if want_disparity else planes_out
