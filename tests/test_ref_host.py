"""Pins the host-side plane-parameter estimation against the REFERENCE'S OWN host code: util::findPeaks
(src/utils/peaks.cpp) and HistogramPeakPlaneParameterProvider::updatePlaneParameters (planeseg.cu:404-458), compiled
verbatim with g++ behind oracle/ref/shim_host.h into oracle/_ref/libref_host.so (oracle/ref/build_ref.sh).  Checked:
the oracle (orc_find_peaks / orc_histogram_peak_update) and the product's cartb200_histogram_peak_update, bit for bit,
on random, plateau-ridden and real derivative histograms."""
import ctypes as C
import os

import numpy as np
import pytest

import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_host.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libref_host.so not built (needs /root/reference at build time)")
    return C.CDLL(LIB)


def histograms():
    rng = np.random.default_rng(99)
    out = []
    for k in range(60):
        kind = k % 6
        if kind == 0:    # two clean bumps like a road + obstacle scene
            x = np.arange(256)
            a, b = rng.integers(100, 126), rng.integers(130, 180)
            h = 4000 * np.exp(-0.5 * ((x - a) / rng.uniform(1, 4)) ** 2) + 2500 * np.exp(-0.5 * ((x - b) / rng.uniform(2, 9)) ** 2)
            h = h + rng.integers(0, 40, 256)
        elif kind == 1:  # heavy ties and plateaus: the unstable std::sort order matters
            h = rng.integers(0, 4, 256) * 100
        elif kind == 2:  # sparse
            h = np.zeros(256)
            h[rng.integers(0, 256, 7)] = rng.integers(1, 1000, 7)
        elif kind == 3:  # monotone / single peak: fewer than two peaks -> parameters stay
            h = np.arange(256) if k % 2 else np.full(256, 5)
        elif kind == 4:  # random
            h = rng.integers(0, 100000, 256)
        else:            # equal neighbouring peaks (zero distance / zero derivative early returns)
            h = np.zeros(256)
            c = int(rng.integers(10, 240))
            h[c], h[c + 1], h[c + 2] = 50, 49, 50
        out.append(np.ascontiguousarray(h, dtype=np.int32))
    gold = os.path.join(HERE, "golden", "ref_kernels.npz")
    if os.path.exists(gold):
        g = np.load(gold)
        for key in g.files:
            if key.startswith("naive_hist_"):  # real derivative histograms (256 bins) from the reference's kernels
                out.append(np.ascontiguousarray(g[key], dtype=np.int32))
    return out


def test_find_peaks_matches_the_reference(ref):
    for h in histograms():
        out = np.zeros((len(h), 4), np.int32)
        n = ref.ref_find_peaks(h.ctypes.data_as(C.c_void_p), len(h), out.ctypes.data_as(C.c_void_p))
        assert np.array_equal(po.find_peaks(h), out[:n])


def test_histogram_peak_update_matches_the_reference(ref):
    import cart_slam_b200 as cb

    changed = 0
    for h in histograms():
        for start in ([0, 0, 0, 0, 0, 0], [7, -2, 1, 30, -3, 1]):
            p = np.array(start, np.int32)
            ref.ref_histogram_peak_update(h.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p))
            _, o = po.histogram_peak_update(h, start)
            assert list(p) == o, (start, list(p), o)
            assert list(p) == list(cb.histogram_peak_update(h, start)[1]), "product (plane_params.cpp) differs from the reference"
            changed += list(p) != start
    assert changed > 20  # the update path, not only the early returns, is exercised


def test_kitti_calibration_matches_the_reference_parser(ref, tmp_path):
    """The host layer's KITTIDataSource (calib.txt -> Q) against the reference's own readLine (kitti.cpp:32-85) on several
    calibration files: the KITTI odometry layout, other numbers / exponents, reordered lines, junk lines."""
    cv2 = pytest.importorskip("cv2")
    from cart_slam_b200 import host

    rng = np.random.default_rng(5)

    def calib_text(k):
        rows = {}
        for cam in range(4):
            fx = 700 + 30 * k + rng.random()
            cx, cy = 600 + rng.random() * 10, 180 + rng.random() * 10
            tx = 0.0 if cam == 0 else float(rng.normal(0, 200))
            v = [fx, 0, cx, tx, 0, fx, cy, rng.normal(), 0, 0, 1, rng.normal() * 1e-3]
            fmt = "%.12e" if k % 2 == 0 else "%.6f"
            rows[cam] = f"P{cam}: " + " ".join(fmt % x for x in v)
        order = [0, 1, 2, 3] if k < 3 else [3, 1, 2, 0]
        lines = [rows[c] for c in order] + ["Tr: " + " ".join("%.6e" % x for x in rng.normal(size=12))]
        if k == 4:
            lines = ["# comment without separator", "Q9 1 2 3"] + lines
        return "\n".join(lines) + "\n"

    f = ref.ref_kitti_q
    f.argtypes = [C.c_char_p, C.c_float, C.c_float, C.c_void_p]
    for k in range(6):
        text = calib_text(k)
        d = tmp_path / f"c{k}" / "sequences" / "00"
        (d / "image_2").mkdir(parents=True)
        (d / "image_3").mkdir()
        (d / "calib.txt").write_text(text)
        cv2.imwrite(str(d / "image_2" / "000000.png"), np.zeros((12, 16, 3), np.uint8))
        cv2.imwrite(str(d / "image_3" / "000000.png"), np.zeros((12, 16, 3), np.uint8))
        W, H, Q = host.open_source({"type": "kitti", "path": str(tmp_path / f"c{k}"), "sequence": 0})
        q = np.zeros(16, np.float32)
        assert f(text.encode(), 1.0, 1.0, q.ctypes.data_as(C.c_void_p)) == 0
        assert (W, H) == (16, 12)
        assert np.array_equal(np.asarray(Q, np.float32).reshape(16), q), k
    # a file without the right camera line
    assert f(b"P2: 1 0 1 1 0 1 1 0 0 0 1 0\n", 1.0, 1.0, np.zeros(16, np.float32).ctypes.data_as(C.c_void_p)) == -1
