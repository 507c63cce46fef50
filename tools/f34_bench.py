"""CUDA-event times of the SURVEY 8(f) f3/f4 kernels at KITTI size (1242x375), one frame per call as the module
surface runs them.  Writes gpurun_out/f34_stage_times.json.  Usage: python tools/f34_bench.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cart_slam_b200 as cb  # noqa: E402


def timed(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3  # us


def main():
    W, H = 1242, 375
    rng = np.random.default_rng(1)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    deriv2 = dev(rng.integers(-8, 40, (H, W, 2)).astype(np.int16))
    bw = (W + 11) // 12
    ys, xs = np.mgrid[0:H, 0:W]
    labels_np = ((ys // 12) * bw + xs // 12).astype(np.uint16)
    n_labels = int(labels_np.max()) + 1
    labels = dev(labels_np)
    pp = [dev(rng.integers(0, 3, (H, W)).astype(np.uint8)) for _ in range(3)]
    pf = [dev((rng.integers(-16, 17, (H, W, 2)) + np.array([96, -16])).astype(np.int16)) for _ in range(3)]
    xyz = dev(rng.uniform(0.5, 60.0, (H, W, 3)).astype(np.float32))
    planes = rng.normal(size=(7, 4))
    out = {}
    with cb.Context(cb.Config(W, H, max_batch=1, enable_sgm=False, sp_block_size=12)) as ctx:
        px = W * H
        out["classify_temporal_3refs_us"] = timed(lambda: ctx.classify_temporal(deriv2, (1, 30, -3, 1), pp, pf))
        out["classify_temporal_bytes"] = px * (4 + 2 + 3 * 5)
        out["sp_planeseg_temporal_3refs_us"] = timed(lambda: ctx.sp_planeseg_temporal(deriv2, labels, (1, 30, -3, 1), pp, pf, max_label=n_labels))
        out["sp_planeseg_temporal_bytes"] = px * (4 + 2 + 1 + 3 * 5 + 2 + 1)
        out["label_statistics_us"] = timed(lambda: ctx.label_statistics(labels, xyz, n_labels))
        out["label_statistics_bytes"] = px * (2 + 4)
        out["region_inliers_7planes_us"] = timed(lambda: ctx.region_inliers(labels, xyz, n_labels, planes, 0.5))
        out["region_inliers_bytes"] = px * (2 + 12)
    for k in ("classify_temporal", "sp_planeseg_temporal", "label_statistics", "region_inliers"):
        us = [v for kk, v in out.items() if kk.startswith(k) and kk.endswith("_us")][0]
        out[k + "_GBps"] = out[k + "_bytes"] / us / 1e3
    out["note"] = "includes the Python/ctypes call overhead of one call per frame; one frame is 0.47 Mpx, so these are launch-latency sized"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "f34_stage_times.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
