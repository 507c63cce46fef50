"""Sweep of the SGM slice size (CARTB200_SGM_SLICE): time of cartb200_disparity for one batch, and bit-identity of the
result with the unsliced run.   python tools/sgm_slice_sweep.py [--batch 64] [--slices 0,8,16,32]"""
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
from stage_bench import timeit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--slices", default="0,8,16,32")
    ap.add_argument("--width", type=int, default=1242)
    ap.add_argument("--height", type=int, default=375)
    ap.add_argument("--disp", type=int, default=128)
    ap.add_argument("--paths", type=int, default=4)
    args = ap.parse_args()
    W, H, D, B = args.width, args.height, args.disp, args.batch
    seq = SyntheticSequence(W, H, D, n_frames=4, tint=True)
    fr = [seq.frame(1 + (i % 4)) for i in range(B)]
    L = torch.from_numpy(np.stack([f[0] for f in fr])).cuda()
    R = torch.from_numpy(np.stack([f[1] for f in fr])).cuda()
    cfg = cb.Config(W, H, max_batch=B, num_disparities=D, paths=args.paths, smoothing_radius=2, smoothing_iterations=1,
                    enable_superpixels=0)
    out = {}
    with cb.Context(cfg) as ctx:
        os.environ["CARTB200_SGM_SLICE"] = "0"
        ref = ctx.disparity(L, R).clone()
        for sl in [int(v) for v in args.slices.split(",")]:
            os.environ["CARTB200_SGM_SLICE"] = str(sl)
            same = bool(torch.equal(ctx.disparity(L, R), ref))
            ms = timeit(lambda: ctx.disparity(L, R), reps=8)
            out[sl] = {"ms": ms, "us_per_frame": ms / B * 1e3, "identical": same}
            print(f"slice {sl:3d}: {ms:8.3f} ms per {B} frames = {ms / B * 1e3:6.1f} us/frame, identical to unsliced: {same}", flush=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sgm_slice_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
