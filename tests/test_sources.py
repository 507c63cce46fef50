"""On-disk KITTI source of the host layer: the PNG reader that replaces cv::imread,
calib.txt -> reprojection matrix Q (kitti.cpp:32-148), the config/sources JSON schema, and - on a GPU - a module
pipeline fed from disk giving the same results as the same frames fed from memory."""
import os

import numpy as np
import pytest

from cart_slam_b200 import host
from cart_slam_b200.synth import SyntheticSequence

cv2 = pytest.importorskip("cv2")

CALIB = """P0: 7.188560000000e+02 0.000000000000e+00 6.071928000000e+02 0.000000000000e+00 0.000000000000e+00 7.188560000000e+02 1.852157000000e+02 0.000000000000e+00 0.000000000000e+00 0.000000000000e+00 1.000000000000e+00 0.000000000000e+00
P1: 7.188560000000e+02 0.000000000000e+00 6.071928000000e+02 -3.861448000000e+02 0.000000000000e+00 7.188560000000e+02 1.852157000000e+02 0.000000000000e+00 0.000000000000e+00 0.000000000000e+00 1.000000000000e+00 0.000000000000e+00
P2: 7.188560000000e+02 0.000000000000e+00 6.071928000000e+02 4.538225000000e+01 0.000000000000e+00 7.188560000000e+02 1.852157000000e+02 -1.130887000000e-01 0.000000000000e+00 0.000000000000e+00 1.000000000000e+00 3.779761000000e-03
P3: 7.188560000000e+02 0.000000000000e+00 6.071928000000e+02 -3.372877000000e+02 0.000000000000e+00 7.188560000000e+02 1.852157000000e+02 2.369057000000e+00 0.000000000000e+00 0.000000000000e+00 1.000000000000e+00 4.915215000000e-03
Tr: 4.276802385584e-04 -9.999672484946e-01 -8.084491683471e-03 -1.198459927713e-02 -7.210626507497e-03 8.081198471645e-03 -9.999413164504e-01 -5.403984729748e-02 9.999738645903e-01 4.859485810390e-04 -7.206933692422e-03 -2.921968648686e-01
"""


def _write_sequence(root, frames, seq=3):
    d = os.path.join(root, "sequences", f"{seq:02d}")
    os.makedirs(os.path.join(d, "image_2"))
    os.makedirs(os.path.join(d, "image_3"))
    open(os.path.join(d, "calib.txt"), "w").write(CALIB)
    for i, (l, r) in enumerate(frames):
        cv2.imwrite(os.path.join(d, "image_2", f"{i:06d}.png"), l)
        cv2.imwrite(os.path.join(d, "image_3", f"{i:06d}.png"), r)
    return d


@pytest.mark.parametrize("kind", ["bgr", "gray", "bgra", "noise_bgr"])
def test_png_reader_matches_opencv(tmp_path, kind):
    rng = np.random.default_rng(5)
    H, W = 37, 53
    if kind == "gray":
        img = rng.integers(0, 256, (H, W), dtype=np.uint8)
    elif kind == "bgra":
        img = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    elif kind == "noise_bgr":
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    else:  # smooth content exercises the Sub / Up / Average / Paeth filters the encoder picks
        yy, xx = np.mgrid[0:H, 0:W]
        img = np.stack([(xx * 3 + yy) % 256, (xx + yy * 5) % 256, (xx * yy) % 256], -1).astype(np.uint8)
    p = str(tmp_path / f"{kind}.png")
    assert cv2.imwrite(p, img)
    ours = host.decode_png(p)
    ref = cv2.imread(p)  # IMREAD_COLOR: BGR, alpha dropped, gray replicated
    assert ours.shape == ref.shape and np.array_equal(ours, ref)


def test_png_reader_rejects_bad_files(tmp_path):
    p = tmp_path / "x.png"
    p.write_bytes(b"not a png at all, definitely not")
    with pytest.raises(host.HostError, match="not a PNG"):
        host.decode_png(str(p))
    with pytest.raises(host.HostError, match="Failed to open image"):
        host.decode_png(str(tmp_path / "missing.png"))
    img16 = np.zeros((4, 4), np.uint16)
    cv2.imwrite(str(tmp_path / "deep.png"), img16)
    with pytest.raises(host.HostError, match="8-bit"):
        host.decode_png(str(tmp_path / "deep.png"))


def test_kitti_source_reads_calibration_like_the_reference(tmp_path):
    frames = [(np.zeros((20, 32, 3), np.uint8), np.zeros((20, 32, 3), np.uint8))]
    _write_sequence(str(tmp_path), frames, seq=3)
    W, H, Q = host.open_source({"type": "kitti", "path": str(tmp_path), "sequence": 3})
    assert (W, H) == (32, 20)
    # kitti.cpp:75-84,141-148 with float arithmetic: baseline = -P[3] / fx of the camera's own projection row
    fx, cx, cy = np.float32(718.856), np.float32(607.1928), np.float32(185.2157)
    bl = -np.float32(45.38225) / fx
    exp = np.eye(4, dtype=np.float32)
    exp[0, 3], exp[1, 3], exp[2, 2], exp[2, 3] = -cx, -cy, 0, fx
    exp[3, 2] = np.float32(-1.0 / np.float64(bl))
    exp[3, 3] = (cx - cx) / bl
    assert np.allclose(Q, exp, rtol=1e-6, atol=0)
    with pytest.raises(host.HostError, match="Failed to open calibration file"):
        host.open_source({"type": "kitti", "path": str(tmp_path), "sequence": 4})
    with pytest.raises(host.HostError, match="Unknown data source type"):
        host.open_source({"type": "bag", "path": "x"})
    with pytest.raises(host.HostError, match="ZED SDK"):
        host.open_source({"type": "zed", "path": "x.svo"})


@pytest.mark.gpu
def test_pipeline_from_disk_equals_pipeline_from_memory(tmp_path):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    W, H, D, n = 192, 96, 64, 4
    seq = SyntheticSequence(W, H, D, n_frames=n, tint=True)
    fr = [seq.frame(i + 1)[:2] for i in range(n)]
    _write_sequence(str(tmp_path), fr, seq=0)
    modules = [{"type": "disparity", "num_disparities": D, "smoothing_radius": 2, "smoothing_iterations": 1},
               {"type": "depth"},
               {"type": "disparity_planeseg", "parameter_provider": {"type": "static", "horizontal_range_min": 1,
                                                                      "horizontal_range_max": 30, "vertical_range_min": -3,
                                                                      "vertical_range_max": 1}}]
    disk = host.run_source({"type": "kitti", "path": str(tmp_path), "sequence": 0}, modules, max_frames=10,
                           want_disparity=True, want_depth=True)
    assert disk["n"] == n
    L = np.stack([f[0] for f in fr])
    R = np.stack([f[1] for f in fr])
    mem = host.run_config(modules, L, R, Q=disk["Q"], want_disparity=True, want_depth=True)
    assert np.array_equal(disk["disparity"], mem["disparity"])
    assert np.array_equal(disk["planes"], mem["planes"])
    assert np.array_equal(disk["depth"], mem["depth"], equal_nan=True)


def test_source_config_image_size_scales_q(tmp_path):
    """"width" / "height" of a kitti source config = KITTIDataSource's imageSize argument: Q is scaled (kitti.cpp:137-148)."""
    seq = SyntheticSequence(64, 32, 64, n_frames=1)
    _write_sequence(str(tmp_path), [seq.frame(1)[:2]], seq=0)
    native = host.open_source({"type": "kitti", "path": str(tmp_path), "sequence": 0})
    scaled = host.open_source({"type": "kitti", "path": str(tmp_path), "sequence": 0, "width": 48, "height": 16})
    assert native[:2] == (64, 32) and scaled[:2] == (48, 16)
    assert np.isclose(scaled[2][0, 3], native[2][0, 3] * 48 / 64) and np.isclose(scaled[2][1, 3], native[2][1, 3] * 16 / 32)
    assert np.isclose(scaled[2][2, 3], native[2][2, 3] * 48 / 64)


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(96, 48), (250, 131), (192, 96)])
def test_resize_kernel_equals_oracle(size):
    """cartb200_resize_bgr8 (the source's cv::cuda::resize replacement) against the oracle's restatement, bit for bit."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import cart_slam_b200 as cb
    import pyoracle as po

    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (96, 192, 3), dtype=np.uint8)
    got = cb.resize_bgr8(torch.from_numpy(img).cuda(), size[0], size[1]).cpu().numpy()
    assert np.array_equal(got, po.resize_bgr8(img, size[0], size[1]))


@pytest.mark.gpu
def test_kitti_source_resizes_on_the_device(tmp_path):
    """A source configured with another image size than the files' (KITTIDataSource's imageSize argument, kitti.cpp:166-169):
    the modules see the resized frames, Q is scaled (kitti.cpp:137-148) - same planes as the pre-resized frames fed from
    memory."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import pyoracle as po

    W, H, D, n = 256, 128, 64, 3
    seq = SyntheticSequence(W, H, D, n_frames=n, tint=True)
    fr = [seq.frame(i + 1)[:2] for i in range(n)]
    _write_sequence(str(tmp_path), fr, seq=0)
    w2, h2 = 192, 96
    modules = [{"type": "disparity", "num_disparities": D, "smoothing_radius": 2, "smoothing_iterations": 1},
               {"type": "disparity_planeseg", "parameter_provider": {"type": "static", "horizontal_range_min": 1,
                                                                      "horizontal_range_max": 30, "vertical_range_min": -3,
                                                                      "vertical_range_max": 1}}]
    disk = host.run_source({"type": "kitti", "path": str(tmp_path), "sequence": 0, "width": w2, "height": h2}, modules,
                           max_frames=10, want_disparity=True)
    assert disk["n"] == n and disk["planes"].shape[1:] == (h2, w2)
    L = np.stack([po.resize_bgr8(f[0], w2, h2) for f in fr])
    R = np.stack([po.resize_bgr8(f[1], w2, h2) for f in fr])
    mem = host.run_config(modules, L, R, want_disparity=True)
    assert np.array_equal(disk["disparity"], mem["disparity"]) and np.array_equal(disk["planes"], mem["planes"])
    native = host.open_source({"type": "kitti", "path": str(tmp_path), "sequence": 0})
    scaled = host.open_source({"type": "kitti", "path": str(tmp_path), "sequence": 0, "width": w2, "height": h2})
    assert native[:2] == (W, H) and scaled[:2] == (w2, h2)
    assert np.isclose(scaled[2][0, 3], native[2][0, 3] * w2 / W) and np.isclose(scaled[2][1, 3], native[2][1, 3] * h2 / H)


def test_own_inflate_equals_zlib_on_every_block_type():
    """The PNG reader's zlib-stream decoder (cart/inflate.hpp) against zlib: stored, fixed and dynamic blocks, every
    level / strategy / window, empty and incompressible inputs; malformed streams are rejected, never mis-decoded."""
    import zlib

    rng = np.random.default_rng(7)
    payloads = [b"", b"x", b"abc" * 2000, bytes(70000), rng.integers(0, 256, 120000, dtype=np.uint8).tobytes(),
                rng.integers(0, 3, 150000, dtype=np.uint8).tobytes(),
                (np.cumsum(rng.integers(-2, 3, 200000)) % 256).astype(np.uint8).tobytes()]
    for data in payloads:
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED):
                c = zlib.compressobj(level=level, strategy=strategy, wbits=15 if level != 1 else 10)
                stream = c.compress(data) + c.flush()
                assert host.inflate(stream, len(data)).tobytes() == data, (len(data), level, strategy)
    data = payloads[-1]
    stream = zlib.compress(data, 6)
    for bad, size in ((stream[:-1], len(data)), (stream[:len(stream) // 2], len(data)), (stream, len(data) - 1),
                      (stream, len(data) + 1), (b"\x78\x9c", 0), (b"", 0), (stream[:-4] + b"\0\0\0\0", len(data)),
                      (b"\x78\x9c" + b"\x07" * 40, 100)):
        with pytest.raises(host.HostError):
            host.inflate(bad, size)
    # random corruption: the verdict and, when accepted, the bytes equal zlib's
    for k in range(300):
        s = bytearray(stream)
        for _ in range(1 + k % 3):
            s[int(rng.integers(2, len(s)))] ^= 1 << int(rng.integers(0, 8))
        try:
            want = zlib.decompress(bytes(s))
            want = want if len(want) == len(data) else None
        except zlib.error:
            want = None
        try:
            got = host.inflate(bytes(s), len(data)).tobytes()
        except host.HostError:
            got = None
        assert got == want, k
