"""Per-stage CUDA-event timings of the cartb200 path at KITTI size (run on the GPU box).
   python tools/stage_bench.py [--batch 16] [--paths 4]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cart_slam_b200 as cb  # noqa: E402
from cart_slam_b200.synth import SyntheticSequence  # noqa: E402


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--paths", type=int, default=4)
    ap.add_argument("--width", type=int, default=1242)
    ap.add_argument("--height", type=int, default=375)
    ap.add_argument("--disp", type=int, default=128)
    ap.add_argument("--block", type=int, default=12)
    ap.add_argument("--tag", default="")
    ap.add_argument("--sp-exact", action="store_true", help="accepted for old command lines; ignored (always exact)")
    ap.add_argument("--once", action="store_true", help="run every stage exactly once (for ncu captures)")
    args = ap.parse_args()
    if args.once:
        global timeit
        def timeit(fn, reps=1, warm=0):  # noqa: F811
            fn()
            torch.cuda.synchronize()
            return 1e-6
    W, H, D, B = args.width, args.height, args.disp, args.batch
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    seq = SyntheticSequence(W, H, D, n_frames=4, tint=True)
    fr = [seq.frame(1 + (i % 4)) for i in range(B)]  # distinct frames (the superpixel work depends on the content)
    L = torch.from_numpy(np.stack([f[0] for f in fr])).cuda()
    R = torch.from_numpy(np.stack([f[1] for f in fr])).cuda()
    cfg = cb.Config(W, H, max_batch=B, num_disparities=D, paths=args.paths, smoothing_radius=2, smoothing_iterations=1,
                    sp_block_size=args.block)
    res = {}
    with cb.Context(cfg) as ctx:
        res["gray_census"] = timeit(lambda: ctx.sgm_gray_census(L, R))
        for p in range(args.paths):
            res[f"aggregate_path{p}"] = timeit(lambda: ctx.sgm_aggregate_path(B, p))
        res["aggregate_all_paths"] = timeit(lambda: ctx.sgm_aggregate(B))
        res["wta_post_interp"] = timeit(lambda: ctx.sgm_wta_post(B))
        disp = ctx.sgm_wta_post(B)
        res["derivative"] = timeit(lambda: ctx.derivative(disp))
        res["naive_derivative"] = timeit(lambda: ctx.naive_derivative(disp))
        deriv, _ = ctx.derivative(disp)
        # every repetition starts from the block initialisation, so all repetitions do the same work
        def sp(its):
            ctx.superpixels_reset(B)
            ctx.superpixels_relax(L, deriv, its)
        r0 = timeit(lambda: sp(0), reps=5)
        r8 = timeit(lambda: sp(2 if args.once else 8), reps=5)
        res["sp_relax_0it"] = r0
        res["sp_relax_per_iteration"] = (r8 - r0) / 8
        labels = ctx.superpixels_relax(L, deriv, 8)
        res["sp_planeseg"] = timeit(lambda: ctx.sp_planeseg(deriv, labels, [[1, 30, -3, 1]]))
    vol_bytes = B * W * H * D
    print(f"batch {B}, {W}x{H}, D={D}; ms per launch (per frame us) [algorithmic GB/s, fraction of {peak:.0f}]")
    for k, v in res.items():
        extra = ""
        if k.startswith("aggregate_path"):
            gbs = (vol_bytes + B * 2 * 4 * W * H) / (v * 1e-3) / 1e9
            extra = f"  {gbs:7.0f} GB/s  {gbs / peak:5.3f}"
        if k == "aggregate_all_paths":
            gbs = args.paths * (vol_bytes + B * 2 * 4 * W * H) / (v * 1e-3) / 1e9
            extra = f"  {gbs:7.0f} GB/s  {gbs / peak:5.3f}"
        if k == "wta_post_interp":
            gbs = (args.paths * vol_bytes) / (v * 1e-3) / 1e9
            extra = f"  {gbs:7.0f} GB/s  {gbs / peak:5.3f} (volume reads only)"
        print(f"  {k:26s} {v:9.3f} ms  ({v / B * 1e3:8.1f} us/frame){extra}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"stage_bench{args.tag}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
