"""The SGM oracle is parity-UNPINNED (the reference's cv::cuda::StereoSGM is third-party and not buildable here,
oracle/sgm.cpp header).  What can be checked on a CPU: the specification recovers the synthetic generator's ground-truth
disparity, and it agrees statistically with OpenCV's CPU StereoSGBM (MODE_HH4, same P1/P2/uniqueness; a different
matching cost, so never bit-exact).  These are plausibility anchors, not parity claims."""
import numpy as np
import pytest

import pyoracle as po
from cart_slam_b200.synth import SyntheticSequence


@pytest.fixture(scope="module")
def pair():
    W, H, D = 400, 200, 64
    seq = SyntheticSequence(W, H, D, n_frames=2)
    l, r, gt = seq.frame(1)
    return l, r, gt, D


def test_oracle_sgm_recovers_ground_truth(pair):
    l, r, gt, D = pair
    d = po.sgm_compute(l, r, D)
    valid = d >= 4 * 16
    assert valid.mean() > 0.8
    assert (np.abs(d / 16.0 - gt)[valid] <= 1.0).mean() > 0.9
    # 8 paths: same ground truth, at least as many valid pixels within a small margin
    d8 = po.sgm_compute(l, r, D, paths=8)
    v8 = d8 >= 4 * 16
    assert (np.abs(d8 / 16.0 - gt)[v8] <= 1.0).mean() > 0.9


def test_oracle_sgm_agrees_with_opencv_cpu_sgbm(pair):
    cv2 = pytest.importorskip("cv2")
    l, r, gt, D = pair
    d = po.sgm_compute(l, r, D)
    sg = cv2.StereoSGBM_create(minDisparity=4, numDisparities=D, blockSize=3, P1=10, P2=120, uniquenessRatio=12,
                               mode=cv2.STEREO_SGBM_MODE_HH4)
    c = sg.compute(cv2.cvtColor(l, cv2.COLOR_BGR2GRAY), cv2.cvtColor(r, cv2.COLOR_BGR2GRAY))
    both = (d >= 4 * 16) & (c >= 4 * 16)
    assert both.mean() > 0.6
    assert (np.abs(d[both] / 16.0 - c[both] / 16.0) <= 1.0).mean() > 0.95
