"""Seeded inputs of the temporal-vote and plane-fit consumer (f4) golden cases, shared by
tests/golden/make_golden_f34.py (which runs the reference's own kernels on them on a GPU box) and the tests."""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))


def make(base):
    """base: the loaded tests/golden/ref_kernels.npz (inputs of the earlier stages are reused)."""
    import pyoracle as po

    rng = np.random.default_rng(3434)
    deriv2 = base["oracle_deriv_smooth"]          # [H, W, 2] int16
    naive = base["naive_deriv_smooth"]            # [H, W] int16
    labels = base["relax_4"]                      # [H, W] uint16
    H, W = labels.shape
    n_labels = int(po.block_init(W, H, 12, 12)[1])
    params = (1, 30, -3, 1)
    # previous frames: the range-rule class of shifted derivatives with 6 % random classes mixed in
    prev_planes, prev_flow = [], []
    for k in range(4):
        pl = po.classify(np.roll(deriv2, (2 * k + 1, -3 * (k + 1)), axis=(0, 1)), *params, channel_stride=2)
        noise = rng.random((H, W)) < 0.06
        pl = np.where(noise, rng.integers(0, 3, (H, W)), pl).astype(np.uint8)
        prev_planes.append(np.ascontiguousarray(pl))
        # S10.5 flow: a camera-like drift of about +3 / -1 px with fractional parts, block-wise outliers (both signs,
        # some larger than the image) and per-pixel jitter across the x.5 boundaries (exercises the arithmetic shift)
        fx = np.full((H, W), 3 * 32 + 7 * k, np.int32) + rng.integers(-40, 41, (H, W))
        fy = np.full((H, W), -1 * 32 - 5 * k, np.int32) + rng.integers(-40, 41, (H, W))
        by, bx = rng.integers(0, H // 16, 40), rng.integers(0, W // 16, 40)
        for j in range(40):
            sl = (slice(by[j] * 16, by[j] * 16 + 16), slice(bx[j] * 16, bx[j] * 16 + 16))
            fx[sl] += int(rng.integers(-W * 40, W * 40))
            fy[sl] += int(rng.integers(-H * 40, H * 40))
        fl = np.stack([np.clip(fx, -32768, 32767), np.clip(fy, -32768, 32767)], axis=2).astype(np.int16)
        prev_flow.append(np.ascontiguousarray(fl))
    # depth: reprojection of the smoothed disparity with a KITTI-like Q (invalid disparities give negative / huge Z,
    # zero disparity gives inf) plus explicit NaN / inf / boundary values
    Q = np.array([[1, 0, 0, -W / 2], [0, 1, 0, -H / 2], [0, 0, 0, 721.5], [0, 0, 1 / 0.54, 0]], np.float32)
    disp = base["smooth_disparity"].copy()
    disp[10:14, 20:60] = 0
    xyz = po.depth(disp, Q)
    xyz[50:54, 100:140, 2] = np.nan
    xyz[60:62, 10:30, 2] = 40.0
    xyz[62:64, 10:30, 2] = np.nextafter(np.float32(40.0), np.float32(50.0))
    xyz[64:66, 10:30, 2] = 0.0
    planes = np.array([[0.0, 1.0, 0.0, 3.0], [0.0, 0.9962, -0.0872, 5.5], [0.02, 0.98, 0.05, 0.3], [1.0, 0.0, 0.0, 2.0],
                       [0.0, 0.0, 1.0, -32.0], [0.1, 0.2, -0.97, 28.0], [0.3, -0.7, 0.2, 1.0]], np.float64)
    d = dict(deriv2=deriv2, naive=naive, labels=labels, n_labels=n_labels, params=params, prev_planes=prev_planes,
             prev_flow=prev_flow, xyz=np.ascontiguousarray(xyz), planes=planes, thresholds=(0.25, 1.5))
    crc = 0
    for a in [deriv2, naive, labels, xyz, planes] + prev_planes + prev_flow:
        crc = zlib.crc32(np.ascontiguousarray(a).tobytes(), crc)
    d["crc"] = crc
    return d
